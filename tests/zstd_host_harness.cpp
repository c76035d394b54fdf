// Host build of strawboat_b200/csrc/sb_zstd.cuh with ONE emulated lane: lets the CPU test suite check the device
// Zstandard decoder against libzstd-written frames without a GPU (tests/test_zstd.py).  Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#define SB_ZSTD_HOST_TEST
#define SB_ZSTD_LANES 1u
#define __device__
#define __forceinline__ inline
#define __constant__ static const
enum { SB_EXTERNAL = 3 };
struct { unsigned x; } threadIdx = {0};
static inline int __clz(int v) { return v ? __builtin_clz(unsigned(v)) : 32; }
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline void __syncwarp() {}
using std::min;
#include "../strawboat_b200/csrc/sb_zstd.cuh"

extern "C" int zstd_harness_decode(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t dlen) {
  static sb::ZstdTables T;
  static uint8_t lits[(128 << 10) + 64];
  return sb::zstd_decode_warp(src, clen, dst, dlen, &T, lits);
}
