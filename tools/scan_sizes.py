"""Developer probe: resident-path kernel times vs batch size (config-2 index)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gramtools_b200 import QuasimapIndex  # noqa: E402

sizes = [int(x) for x in os.environ.get("GQ_SIZES", "131072,262144,524288,1000000,2000000").split(",")]
prg, bases, offs, seeds = bench.make_workload(0, max(sizes))
idx = QuasimapIndex(prg, bench.KMER)
for kv in os.environ.get("GQ_OPTIONS", "").split(","):
    if "=" in kv:
        k, v = kv.split("=")
        idx.set_option(k, int(v))
for n in sizes:
    idx.upload(bases[:int(offs[n])], offs[:n + 1], seeds[:n])
    best = None
    for _ in range(5):
        idx.map_resident()
        i = idx.run_info()
        t = (i["search_ms"], i["coverage_ms"], i["kernels_ms"])
        best = t if best is None or t[2] < best[2] else best
    print(f"n_reads={n}: search {best[0]:.3f} ms, classify+coverage {best[1]:.3f} ms, all kernels {best[2]:.3f} ms -> "
          f"{n / (best[2] / 1e3) / 1e6:.0f} M reads/s (kernels)", flush=True)
