#!/bin/bash
# per-launch kernel times of one resident config-2 batch (ncu serialises and cold-caches: compare shares)
out=${1:-gpurun_out/launches.csv}
GQ_PROFILE_ITERS=${GQ_PROFILE_ITERS:-2} ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out python tools/profile_run.py > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:]:
    print(r[ki].split("(")[0][-40:].ljust(42), float(r[vi].replace(",", "")) / 1e3, "us")
PY
