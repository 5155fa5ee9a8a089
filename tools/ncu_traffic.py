"""profiles/r02_dram_traffic.json from an `ncu --page raw --csv` export: DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) and duration per kernel of one step, per read. bench.py reports them beside the byte model.

  python tools/ncu_traffic.py <raw.csv> <config index> <reads in the profiled batch> [source tag]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, config, n_reads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tag = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(raw)
rows = list(csv.reader(open(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    x = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


per_kernel, ms = {}, {}
for r in body:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    per_kernel[name] = per_kernel.get(name, 0.0) + val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    ms[name] = ms.get(name, 0.0) + t * {"us": 1e-3, "ms": 1.0, "ns": 1e-6}.get(units[col["gpu__time_duration.sum"]].lower(), 1e-3)
out_path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[f"config{config}"] = {
    "source": f"ncu --set full --clock-control none ({tag}); per-launch values, cold caches, kernels serialised",
    "reads": n_reads,
    "dram_bytes_per_read": {k: v / n_reads for k, v in per_kernel.items()},
    "ncu_ms": ms,
}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data[f"config{config}"], indent=1))
