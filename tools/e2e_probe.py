"""Developer probe: end-to-end time of one batch from pinned host buffers (gq_map_batch_packed / gq_map_batch) vs
pipeline slice sizes, and the pieces around it (reset, coverage fetch). GQ_CHUNKS = comma list of
chunk_reads[:tail_chunk_reads]."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gramtools_b200 import QuasimapIndex, pack_reads  # noqa: E402

prg, bases, offs, seeds = bench.make_workload(0, 1_000_000)
idx = QuasimapIndex(prg, bench.KMER)
pin = lambda a, t: torch.from_numpy(a.view(t)).pin_memory().numpy().view(a.dtype)
pb, po, ps = pin(bases, np.uint8), pin(offs, np.int64), pin(seeds, np.int32)
pk, pw, pl = (pin(x, np.int32) for x in pack_reads(bases, offs))
idx.upload(pb, po, ps)
for _ in range(3):
    idx.map_resident()
print("resident:", idx.run_info())


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


print(f"reset_coverage {timed(idx.reset_coverage):.3f} ms; coverage fetch {timed(idx.coverage):.3f} ms")
for kv in os.environ.get("GQ_OPTIONS", "").split(","):
    if "=" in kv:
        k_, v_ = kv.split("=")
        idx.set_option(k_, int(v_))
for spec in os.environ.get("GQ_CHUNKS", "1048576:1048576,524288:65536,262144:32768,131072:32768,65536:32768").split(","):
    chunk, tail = (int(x) for x in (spec.split(":") + [spec])[:2])
    idx.set_option("chunk_reads", chunk)
    idx.set_option("tail_chunk_reads", tail)
    dp = timed(lambda: idx.map_batch_packed(pk, pw, pl, ps))
    info = idx.run_info()
    du = timed(lambda: idx.map_batch(pb, po, ps))
    print(f"chunk_reads={chunk} tail={tail}: packed {dp:.3f} ms ({1e3/dp:.0f} M reads/s), kernels+waits on the stream "
          f"{info['kernels_ms']:.3f} ms, host enqueue {info['enqueue_ms']:.2f} ms, {info['launches']} launches; u8 {du:.3f} ms")
