// sb_nested.cuh -- rep/def level section of nested pages (read_validity_nested,
// src/read/read_basic.rs:65-173; level streams: SURVEY App. D.3; Dremel walk: App. D.4).
//
// Page layout: [u32 rows][u32 rep_len][u32 def_len][rep stream][def stream][VALUE_BLOCK].
// The reference walks the (rep, def) pairs serially and pushes (child length, validity)
// into one `Nested` per depth.  Every push decision depends only on the entry's own
// (rep, def) and on static per-depth thresholds, and the pushed child length is "how many
// pushes did depth d+1 see so far" -- so the walk is a set of per-depth stream compactions:
//   flags  : pushed_d(e), valid_d(e)                    (independent per entry)
//   scan   : pos_d(e) = #pushes at depth d before e      (block scans, chunk by chunk)
//   write  : offsets_d[pos_d] = base_{d+1} + pos_{d+1}(e);  validity_d bit pos_d = valid_d
// Pass 0 only counts pushes per depth (the page's contribution to every depth's length);
// the host turns the counts into bases; pass 1 writes.
#pragma once
#include "sb_decode.cuh"

namespace sb {

// ------------------------------------------------------------------------------------
// hybrid-RLE level stream -> one byte per entry (parquet2 HybridRleDecoder semantics:
// bit-packed runs may have a short tail; RLE runs carry ceil(w/8) value bytes).
// Thread 0 walks the run headers in batches; all threads expand the runs of a batch.
// ------------------------------------------------------------------------------------
struct LvlRun {
  uint32_t start, count; // entries [start, start+count)
  uint32_t pos;          // bit-packed: byte position of the packed data in the stream
  uint32_t value;        // RLE: the value; bit-packed: 0xffffffff
};
constexpr uint32_t kLvlRuns = 32;

__device__ bool levels_decode(Dctx &cx, const uint8_t *s, uint32_t len, uint32_t w, uint32_t n, uint8_t *out, LvlRun *runs) {
  const uint32_t tid = threadIdx.x;
  if (w == 0) {
    for (uint32_t i = tid; i < n; i += SB_NT) out[i] = 0;
    return true;
  }
  if (w > 8) { // levels wider than 8 bits need > 255 nesting depths
    cx.flag(SB_NYI);
    return false;
  }
  uint32_t done = 0, pos = 0;
  while (done < n) {
    __syncthreads();
    if (tid == 0) {
      uint32_t nr = 0, d = done, p = pos;
      int rc = 0;
      while (d < n && nr < kLvlRuns) {
        if (p >= len) {
          rc = SB_OUT_OF_SPEC;
          break;
        }
        uint64_t header = 0;
        uint32_t used = 0;
        for (uint32_t i = 0; i < 10 && p + i < len; ++i) {
          uint32_t b = s[p + i];
          header |= uint64_t(b & 0x7f) << (7 * i);
          if (!(b & 0x80)) {
            used = i + 1;
            break;
          }
        }
        if (!used) {
          rc = SB_OUT_OF_SPEC;
          break;
        }
        p += used;
        LvlRun r;
        r.start = d;
        if (header & 1) {
          uint64_t bytes = (header >> 1) * w;
          if (bytes > len - p) bytes = len - p;
          uint64_t avail = bytes * 8 / w;
          uint32_t take = uint32_t(avail < uint64_t(n - d) ? avail : uint64_t(n - d));
          if (take == 0) {
            rc = SB_OUT_OF_SPEC;
            break;
          }
          r.count = take;
          r.pos = p;
          r.value = 0xffffffffu;
          p += uint32_t(bytes);
        } else {
          uint64_t run = header >> 1;
          if (run == 0 || p + 1 > len) {
            rc = SB_OUT_OF_SPEC;
            break;
          }
          r.value = s[p];
          r.pos = 0;
          p += 1;
          r.count = uint32_t(run < uint64_t(n - d) ? run : uint64_t(n - d));
        }
        runs[nr++] = r;
        d += r.count;
      }
      cx.bcast[0] = rc;
      cx.bcast[1] = int(nr);
      cx.bcast[2] = int(d);
      cx.bcast[3] = int(p);
    }
    __syncthreads();
    int rc = cx.bcast[0];
    uint32_t nr = uint32_t(cx.bcast[1]);
    if (rc) {
      cx.flag(rc);
      return false;
    }
    const uint32_t mask = (1u << w) - 1u;
    for (uint32_t r = 0; r < nr; ++r) {
      const LvlRun run = runs[r];
      if (run.value != 0xffffffffu) {
        for (uint32_t i = tid; i < run.count; i += SB_NT) out[run.start + i] = uint8_t(run.value);
      } else {
        const uint8_t *p = s + run.pos;
        for (uint32_t i = tid; i < run.count; i += SB_NT) {
          uint32_t bit = i * w;
          uint32_t v = uint32_t(p[bit >> 3]) >> (bit & 7);
          if ((bit & 7) + w > 8) v |= uint32_t(p[(bit >> 3) + 1]) << (8 - (bit & 7)); // inside `bytes` because take <= avail
          out[run.start + i] = uint8_t(v & mask);
        }
      }
    }
    done = uint32_t(cx.bcast[2]);
    pos = uint32_t(cx.bcast[3]);
  }
  __syncthreads();
  return true;
}

// push / validity masks of one entry: bit d set <=> depth d pushes (is valid)
__device__ __forceinline__ void nest_entry(const ColDesc &ni, uint32_t rep, uint32_t def, uint32_t *push, uint32_t *valid) {
  uint32_t pm = 0, vm = 0;
  bool is_required = false;
#pragma unroll
  for (int d = 0; d < SB_MAX_NESTED; ++d) {
    if (d < ni.n_nested) {
      bool right = rep <= ni.cum_rep[d] && def >= ni.cum_sum[d];
      if (is_required || right) {
        bool v = ni.nnull[d] && def > ni.cum_sum[d];
        pm |= 1u << d;
        if (d == ni.n_nested - 1) { // leaf validity bit (read_basic.rs:139-149)
          bool lv = (def != ni.cum_sum[d]) || !ni.nnull[d];
          if (right && lv) vm |= 1u << d;
        } else if (v) {
          vm |= 1u << d;
        }
        is_required = ni.kind[d] == SB_N_STRUCT && !v;
      }
    }
  }
  *push = pm;
  *valid = vm;
}

// Decodes the level section.  Pass 0: counts[d] = pushes at depth d.  Pass 1: writes the
// NestedState entries / leaf validity.  Returns the byte offset of the value block or
// 0xffffffff; *leaf_len = number of leaf slots (values in the VALUE_BLOCK).
__device__ uint32_t decode_levels(Dctx &cx, const uint8_t *page, uint32_t avail, uint32_t n, const ColDesc &ni, int pass,
                                  PageAux *ax /* global: pass 0 writes cnt[], pass 1 reads base[] */, bool last_page, uint32_t *leaf_len) {
  constexpr uint32_t EPT = 8, CH = SB_NT * EPT;
  const uint32_t tid = threadIdx.x;
  const int D = ni.n_nested;
  if (avail < 12) {
    cx.flag(SB_IO);
    return 0xffffffffu;
  }
  const uint32_t additional = ld_u32u(page), rep_len = ld_u32u(page + 4), def_len = ld_u32u(page + 8);
  if (uint64_t(rep_len) + def_len > uint64_t(avail - 12)) {
    cx.flag(SB_IO);
    return 0xffffffffu;
  }
  const uint32_t vb = 12 + rep_len + def_len;
  const uint32_t max_rep = ni.cum_rep[D], max_def = ni.cum_sum[D];
  const uint32_t w_rep = 32 - __clz(max_rep), w_def = 32 - __clz(max_def);

  Arena mark = cx.ar;
  LvlRun *runs = static_cast<LvlRun *>(cx.ar.alloc(sizeof(LvlRun) * kLvlRuns));
  uint8_t *reps = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(n) + 16));
  uint8_t *defs = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(n) + 16));
  uint32_t *stage[SB_MAX_NESTED];
  bool ok = runs && reps && defs;
  const uint32_t stage_words = (n + 31) / 32 + 1;
#pragma unroll
  for (int d = 0; d < SB_MAX_NESTED; ++d) {
    stage[d] = nullptr;
    if (pass == 1 && d < D && ni.nnull[d] && ok) {
      stage[d] = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(stage_words) * 4));
      ok = ok && stage[d];
    }
  }
  if (!ok) {
    cx.flag(SB_NYI);
    return 0xffffffffu;
  }
  if (!levels_decode(cx, page + 12, rep_len, w_rep, n, reps, runs)) return 0xffffffffu;
  if (!levels_decode(cx, page + 12 + rep_len, def_len, w_def, n, defs, runs)) return 0xffffffffu;
  if (pass == 1) {
#pragma unroll
    for (int d = 0; d < SB_MAX_NESTED; ++d)
      if (stage[d])
        for (uint32_t i = tid; i < stage_words; i += SB_NT) stage[d][i] = 0;
  }

  // ---- how many entries does the reference consume?  It stops after entry e when the next
  //      entry starts a row (rep == 0, or end of stream) and `additional` rows were seen.
  uint32_t *s_lim = reinterpret_cast<uint32_t *>(cx.bcast);
  __syncthreads();
  if (tid == 0) s_lim[0] = n;
  __syncthreads();
  {
    uint32_t zrun = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += CH) {
      uint32_t z = 0;
      const uint32_t e0 = c0 + tid * EPT;
#pragma unroll
      for (uint32_t j = 0; j < EPT; ++j)
        if (e0 + j < n && reps[e0 + j] == 0) ++z;
      uint32_t total;
      uint32_t zpre = zrun + block_excl_scan(z, cx.ws, &total);
#pragma unroll
      for (uint32_t j = 0; j < EPT; ++j) {
        uint32_t e = e0 + j;
        if (e < n) {
          if (reps[e] == 0) ++zpre;
          bool next_zero = (e + 1 >= n) || reps[e + 1] == 0;
          if (next_zero && zpre == additional) atomicMin(s_lim, e + 1);
        }
      }
      zrun += total;
      if (zrun > additional) break; // uniform: the stop point has been passed
    }
  }
  __syncthreads();
  const uint32_t P = s_lim[0];
  __syncthreads();

  // ---- per-depth compaction
  uint32_t run_cnt[SB_MAX_NESTED];
  uint64_t base[SB_MAX_NESTED];
#pragma unroll
  for (int d = 0; d < SB_MAX_NESTED; ++d) {
    run_cnt[d] = 0;
    base[d] = pass == 1 ? ax->base[d] : 0;
  }
  for (uint32_t c0 = 0; c0 < P; c0 += CH) {
    const uint32_t e0 = c0 + tid * EPT;
    uint32_t pm[EPT], vm[EPT];
    uint32_t cnt_lo = 0, cnt_hi = 0; // 8 depths x 8-bit per-thread counts (<= EPT each), 4 per word
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      pm[j] = vm[j] = 0;
      if (e0 + j < P) {
        nest_entry(ni, reps[e0 + j], defs[e0 + j], &pm[j], &vm[j]);
        uint32_t m = pm[j];
        cnt_lo += (m & 1u) | ((m & 2u) << 7) | ((m & 4u) << 14) | ((m & 8u) << 21);
        m >>= 4;
        cnt_hi += (m & 1u) | ((m & 2u) << 7) | ((m & 4u) << 14) | ((m & 8u) << 21);
      }
    }
    uint32_t excl[SB_MAX_NESTED];
#pragma unroll
    for (int d = 0; d < SB_MAX_NESTED; ++d) {
      excl[d] = 0;
      if (d < D) {
        uint32_t c = ((d < 4 ? cnt_lo : cnt_hi) >> (8 * (d & 3))) & 0xffu, total;
        uint32_t ex = block_excl_scan(c, cx.ws, &total);
        excl[d] = run_cnt[d] + ex;
        run_cnt[d] += total;
      }
    }
    if (pass == 1) {
#pragma unroll
      for (uint32_t j = 0; j < EPT; ++j) {
#pragma unroll
        for (int d = 0; d < SB_MAX_NESTED; ++d) {
          if (d < D && (pm[j] >> d) & 1u) {
            const uint32_t pos = excl[d];
            if (ni.kind[d] == SB_N_LIST) {
              // child length so far: depth d+1 is pushed after depth d within the same entry
              uint64_t child = (d + 1 < D) ? base[d + 1] + excl[d + 1] : 0;
              ni.nest_off[d][base[d] + pos] = int64_t(child);
            }
            if (stage[d] && ((vm[j] >> d) & 1u)) atomicOr(&stage[d][pos >> 5], 1u << (pos & 31));
          }
        }
#pragma unroll
        for (int d = 0; d < SB_MAX_NESTED; ++d)
          if (d < D && (pm[j] >> d) & 1u) ++excl[d];
      }
    }
  }
  __syncthreads();
  if (pass == 0) {
    if (tid == 0) {
#pragma unroll
      for (int d = 0; d < SB_MAX_NESTED; ++d) ax->cnt[d] = d < D ? run_cnt[d] : 0u;
    }
  } else {
#pragma unroll
    for (int d = 0; d < SB_MAX_NESTED; ++d) {
      if (stage[d]) {
        BitsPacked bs{reinterpret_cast<const uint8_t *>(stage[d])};
        emit_bits(d == D - 1 ? ni.validity : ni.nest_val[d], base[d], run_cnt[d], bs);
      }
      // arrow2 create_list appends the end offset (= child length) after the last page
      if (last_page && tid == 0 && d < D && ni.kind[d] == SB_N_LIST)
        ni.nest_off[d][base[d] + run_cnt[d]] = int64_t((d + 1 < D) ? base[d + 1] + run_cnt[d + 1] : 0);
    }
    __syncthreads();
  }
  *leaf_len = run_cnt[D - 1];
  cx.ar = mark;
  return vb;
}

} // namespace sb
