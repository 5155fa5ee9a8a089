"""Small nested / SNP / indel batches through every kernel route, for `compute-sanitizer --tool memcheck|racecheck`
(developer aid; parity is asserted too)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import Oracle, assert_parity, gpu_result  # noqa: E402
from gramtools_b200 import QuasimapIndex, master_seeds, pack_reads, synth  # noqa: E402


def reads_for(prg, n, L, seed):
    rng = np.random.default_rng(seed)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    return synth.sample_reads(haps, n, L, seed, frac_garbage=0.05, frac_n=0.02)


cases = [("nested", synth.make_nested_prg(4, 300, 3), 4, 40, {}),
         ("nested-small-arena", synth.make_nested_prg(3, 250, 11), 3, 25, {"arena_words": 64, "gtab_cap": 4}),
         ("snp", synth.make_snp_prg(3000, 200, 5)[0], 6, 70, {}),
         ("indel", synth.make_indel_prg(3000, 150, 5), 6, 60, {"pool_words_per_read": 1})]
for name, prg, k, L, opts in cases:
    bases, offs = reads_for(prg, 1500, L, 7)
    seeds = master_seeds(42, offs.size - 1)
    idx = QuasimapIndex(prg, k)
    for kk, vv in opts.items():
        idx.set_option(kk, vv)
    idx.map_batch(bases, offs, seeds)
    got = gpu_result(idx)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    ref = o.result()
    assert_parity(got, ref, name)
    idx.reset_coverage()
    idx.map_batch_packed(*pack_reads(bases, offs), seeds)
    assert_parity(gpu_result(idx), ref, name + "/packed")
    idx.close()
    print(name, "ok", got.stats, flush=True)
