"""One config-2 batch through the resident path a few times — the command ncu wraps
(see profiles/README.md). Not a benchmark: numbers printed under a profiler are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gramtools_b200 import QuasimapIndex  # noqa: E402

n_reads = int(os.environ.get("GQ_PROFILE_READS", 1_000_000))
iters = int(os.environ.get("GQ_PROFILE_ITERS", 3))
prg, bases, offs, seeds = bench.make_workload(0, n_reads)
idx = QuasimapIndex(prg, bench.KMER)
for kv in os.environ.get("GQ_OPTIONS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        idx.set_option(k, int(v))
idx.upload(bases, offs, seeds)
for _ in range(iters):
    idx.map_resident()
    print(idx.run_info())
