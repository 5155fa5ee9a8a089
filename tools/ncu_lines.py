"""Join an `ncu --page source --csv` (SASS view) export of one kernel with the line table of the cubin
(`nvdisasm -g -c`) so that executed-instruction counts and stall samples can be read per source line.

  cuobjdump -xelf all gramtools_b200/libgq.so          # -> kernels.sm_100a.cubin
  nvdisasm -g -c kernels.sm_100a.cubin > k.sass
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:search_kernel > src.csv
  python tools/ncu_lines.py src.csv k.sass 'search_kernelILb1' [top_n]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, func = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60

# offset -> (file, line) for the chosen function
line_of, cur, in_func = {}, None, False
for ln in open(sass):
    if ln.startswith(".text."):
        in_func = func in ln
        continue
    if not in_func:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
body = []
for r in rows[rows.index(hdr) + 1:]:  # first kernel instance only
    if not r or not r[0].startswith("0x"):
        break
    body.append(r)
ia, ie, it, ismp = (hdr.index(x) for x in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base = int(body[0][ia], 16)
per_line = defaultdict(lambda: [0, 0, 0, 0])
tot_i = tot_s = 0
for r in body:
    off = int(r[ia], 16) - base
    loc, _ = line_of.get(off, (("?", 0), ""))
    e, t, s = int(r[ie]), int(r[it]), int(r[ismp])
    p = per_line[loc]
    p[0] += e
    p[1] += t
    p[2] += s
    p[3] += 1
    tot_i += e
    tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}, sass instructions {len(body)}")
srcs = {}
for (f, l), p in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(f"/root/repo/gramtools_b200/csrc/{f}").read().split("\n")
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{100 * p[0] / tot_i:5.1f}% inst {100 * p[2] / max(tot_s, 1):5.1f}% smp  lanes {p[1] / max(p[0], 1):4.1f}  sass {p[3]:3d}  {f}:{l}  {text}")
