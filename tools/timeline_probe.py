"""Developer probe: GQ_TIMELINE=1 python tools/timeline_probe.py — per-slice copy / kernel times of the pipelined
host path (printed by the library), plus a raw pinned H2D copy timing for comparison. GQ_OPTIONS=name=value,..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gramtools_b200 import QuasimapIndex, pack_reads  # noqa: E402

prg, bases, offs, seeds = bench.make_workload(0, 1_000_000)
idx = QuasimapIndex(prg, bench.KMER)
for kv in os.environ.get("GQ_OPTIONS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        idx.set_option(k, int(v))
pin = lambda a, t: torch.from_numpy(a.view(t)).pin_memory().numpy().view(a.dtype)
ps = pin(seeds, np.int32)
pk, pw, pl = (pin(x, np.int32) for x in pack_reads(bases, offs))
for i in range(4):
    sys.stderr.write(f"--- call {i}\n")
    idx.map_batch_packed(pk, pw, pl, ps)
# raw copies of the same pinned buffer with torch, for comparison
src = torch.from_numpy(pk.view(np.int32))
dst = torch.empty_like(src, device="cuda")
for nbytes in (1 << 20, 4 << 20, 14 << 20, 40 << 20):
    n = nbytes // 4
    for _ in range(3):
        dst[:n].copy_(src[:n], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dst[:n].copy_(src[:n], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sys.stderr.write(f"raw H2D {nbytes >> 20} MB: {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s\n")
