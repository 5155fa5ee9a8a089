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
    print("   kernel ms:", {k: round(v, 4) for k, v in idx.kernel_ms().items()})
if os.environ.get("GQ_DEBUG"):
    import ctypes as C
    import numpy as np
    from gramtools_b200 import load_library
    a = np.zeros(32, dtype=np.uint64)
    load_library().gq_debug_counters(a.ctypes.data_as(C.POINTER(C.c_uint64)))
    a = [int(x) / iters for x in a]
    names = ["outer", "sum_run", "sum_wait", "wide_n", "wide_sum", "scan_n", "scan_sum", "top_n", "top_sum", "pop_n",
             "pop_sum", "refill_n", "refill_sum", "hot_n", "hot_sum_run_after"]
    print({k: v for k, v in zip(names, a)})
    print("avg run at vote %.1f, avg waiting %.1f; batch sizes: wide %.1f scan %.1f top %.1f pop %.1f refill %.1f; hot iters %.0f avg run after %.1f"
          % (a[1] / a[0], a[2] / a[0], a[4] / max(a[3], 1), a[6] / max(a[5], 1), a[8] / max(a[7], 1), a[10] / max(a[9], 1),
             a[12] / max(a[11], 1), a[13], a[14] / max(a[13], 1)))
