"""Developer sweep on the GPU: random PRG shapes / k / read lengths through libgq (C ABI) against the oracle —
the same generator as tests/test_host_parity.py::test_random_shapes, more cases."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import Oracle, assert_parity, gpu_result  # noqa: E402
from gramtools_b200 import QuasimapIndex, master_seeds, synth  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t0, fails = time.time(), 0
for case in range(n_cases):
    rng = np.random.default_rng(1000 + case)
    kind = case % 4
    if kind == 0:
        prg = synth.make_snp_prg(int(rng.integers(300, 4000)), int(rng.integers(10, 300)), case)[0]
    elif kind == 1:
        prg = synth.make_nested_prg(int(rng.integers(1, 6)), int(rng.integers(100, 400)), case)
    elif kind == 2:
        prg = synth.make_indel_prg(int(rng.integers(500, 4000)), int(rng.integers(10, 200)), case)
    else:
        base = rng.integers(1, 5, int(rng.integers(200, 1500))).astype(np.uint32)
        u = base[10:10 + int(rng.integers(30, 150))]
        prg = np.concatenate([base, u, rng.integers(1, 5, 100).astype(np.uint32), u, base[:50]]).astype(np.uint32)
    k = int(rng.integers(2, 8))
    L = int(rng.integers(k, 90))
    hrng = np.random.default_rng(case)
    haps = [synth.random_haplotype(prg, hrng) for _ in range(4)]
    bases, offs = synth.sample_reads(haps, 300, L, case, frac_garbage=0.05, frac_n=0.01)
    if float(rng.choice([0.0, 0.0, 0.01])) > 0:
        bases = bases.copy()
        hit = rng.random(bases.size) < 0.01
        bases[hit] = rng.integers(1, 5, int(hit.sum()))
    seeds = master_seeds(int(rng.integers(0, 1000)), offs.size - 1)
    try:
        idx = QuasimapIndex(prg, k, device=0)
        idx.set_option("arena_words", int(rng.choice([64, 256, 1024])))
        idx.map_batch(bases, offs, seeds)
        got = gpu_result(idx)
        idx.close()
        o = Oracle(prg, k)
        o.map(bases, offs, seeds)
        assert_parity(got, o.result(), f"case{case}")
    except Exception as ex:  # noqa: BLE001
        fails += 1
        print("FAIL case", case, "kind", kind, "k", k, "L", L, repr(ex)[:300], flush=True)
print("cases", n_cases, "fails", fails, "time", round(time.time() - t0, 1))
