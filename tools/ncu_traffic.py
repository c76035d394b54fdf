"""`.ncu-rep` of `bench.py --steps 2 --warmup 3 --no-extras` (ncu --set full) -> profiles/r2_ncu_traffic.json:
dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel of the step, keyed by kernel name, together
with the hash of the kernel sources the capture was taken on.  bench.py reports `roofline.traffic` from this file
only while the hash still matches the sources that are running (a stale figure is worse than none).
usage: python tools/ncu_traffic.py gpurun_out/x.ncu-rep"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}


def val(r, m):
    return float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)


kern = {}
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0]
    k = kern.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "time_ns": 0.0})
    k["launches"] += 1
    k["dram_read"] += val(r, "dram__bytes_read.sum")
    k["dram_write"] += val(r, "dram__bytes_write.sum")
    k["time_ns"] += val(r, "gpu__time_duration.sum")
for k in kern.values():  # per launch
    n = k["launches"]
    k["dram_read"], k["dram_write"], k["time_ns"] = k["dram_read"] / n, k["dram_write"] / n, k["time_ns"] / n
    k["dram_bytes"] = k["dram_read"] + k["dram_write"]
out = {"kernel_sources_sha": bench.kernel_sources_hash(), "capture": os.path.basename(sys.argv[1]),
       "command": "ncu --set full --clock-control none --import-source on -k regex:sb_ python bench.py --steps 2 --warmup 3 --no-extras",
       "note": "per launch, averaged over the captured launches; times are under the profiler (serialised, cold cache)", "kernels": kern}
path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
