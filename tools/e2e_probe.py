"""Developer probe: end-to-end gq_map_batch time (pinned host buffers) vs pipeline slice sizes.
GQ_CHUNKS = comma list of chunk_reads[:tail_chunk_reads]."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gramtools_b200 import QuasimapIndex  # noqa: E402

prg, bases, offs, seeds = bench.make_workload(0, 1_000_000)
idx = QuasimapIndex(prg, bench.KMER)
pb = torch.from_numpy(bases).pin_memory().numpy()
po = torch.from_numpy(offs.view(np.int64)).pin_memory().numpy().view(np.uint64)
ps = torch.from_numpy(seeds.view(np.int32)).pin_memory().numpy().view(np.uint32)
for spec in os.environ.get("GQ_CHUNKS", "131072:131072,262144:32768,262144:65536,524288:32768").split(","):
    chunk, tail = (int(x) for x in (spec.split(":") + [spec])[:2])
    idx.set_option("chunk_reads", chunk)
    idx.set_option("tail_chunk_reads", tail)
    for _ in range(3):
        idx.map_batch(pb, po, ps)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        idx.map_batch(pb, po, ps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    t = time.perf_counter()
    for _ in range(10):
        idx.coverage()
    dc = (time.perf_counter() - t) / 10
    info = idx.run_info()
    print(f"chunk_reads={chunk} tail={tail}: map_batch {dt*1e3:.2f} ms ({1e6/dt/1e6:.0f} M reads/s); coverage fetch {dc*1e3:.2f} ms; "
          f"host enqueue {info['enqueue_ms']:.2f} ms, {info['launches']} launches")
