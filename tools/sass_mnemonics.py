"""Mnemonic counts per kernel from `cuobjdump -sass gramtools_b200/libgq.so` (profiles/*_sass_mnemonics.txt): the
instructions that show what the kernels are made of — 256-bit sector loads (LDG.E.ENL2.256: rank blocks), TMA bulk
copies + mbarrier waits (UBLKCP, SYNCS: rank superblock counters into shared memory), warp votes / reductions /
shuffles (VOTE, REDUX, SHFL: the warp-synchronous loops), fire-and-forget reductions (RED: coverage counters)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gramtools_b200/libgq.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keep = re.compile(r"^(LDG\.E\.ENL2\.256|LDG|UBLKCP|SYNCS|REDUX|RED|ATOMG|ATOMS|LDS|STS|SHFL|VOTE|POPC|BREV|FLO|LDL|STL|BSSY|BSYNC)")
fn, cnt, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(2)
        total[fn] += 1
        k = keep.match(op)
        if k:
            key = "LDG.E.ENL2.256" if op.startswith("LDG.E.ENL2.256") else k.group(1)
            cnt[fn][key] += 1
print("# cuobjdump -sass", lib, "(sm_100a): instruction counts per kernel")
for f in sorted(cnt):
    name = re.sub(r"^_ZN2gq\d+", "", f)
    print(f"{name[:60]:60s} total {total[f]:6d}  " + "  ".join(f"{k} {v}" for k, v in sorted(cnt[f].items())))
