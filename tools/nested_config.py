"""Developer run of the config-3 shape on one GPU: nested-variant PRG (n_loci x locus_len, bracket grammar with
nested / adjacent / empty alleles), 150 bp reads. Checks the first `n_check` reads bit-for-bit against the oracle
and prints kernel times of the whole batch — this is the workload that exercises the general kernel."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import Oracle, assert_parity, gpu_result  # noqa: E402
from gramtools_b200 import QuasimapIndex, synth  # noqa: E402

n_loci, locus_len, k, n_reads, n_check = [int(x) for x in (sys.argv[1:6] + ["200", "5000", "10", "1000000", "20000"][len(sys.argv) - 1:])]
t = time.time()
distinct = os.environ.get("GQ_DISTINCT", "1") != "0"  # 1: the bench's config-3 PRG; 0: coinciding alleles (stress)
prg = synth.make_nested_prg(n_loci, locus_len, 0x6772616D + 3, distinct=distinct)
print(f"PRG: {prg.size} symbols ({time.time() - t:.1f} s)", flush=True)
rng = np.random.default_rng(3)
haps = [synth.random_haplotype(prg, rng) for _ in range(8)]
bases, offs = synth.sample_reads(haps, n_reads, 150, 13)
seeds = synth.master_seeds(42, n_reads)
t = time.time()
idx = QuasimapIndex(prg, k)
print(f"index build+upload {time.time() - t:.1f} s; nested={idx.layout.is_nested} sites={idx.layout.n_sites}", flush=True)
# parity on a prefix
m = min(n_check, n_reads)
idx.map_batch(bases[:int(offs[m])], offs[:m + 1], seeds[:m])
got = gpu_result(idx)
o = Oracle(prg, k)
o.map(bases[:int(offs[m])], offs[:m + 1], seeds[:m], threads=os.cpu_count())
assert_parity(got, o.result(), "nested-config prefix")
print(f"parity ok on {m} reads: stats {got.stats}", flush=True)
idx.reset_coverage()
idx.upload(bases, offs, seeds)
for it in range(3):
    idx.map_resident()
    info = idx.run_info()
    print(info, f"-> {n_reads / (info['kernels_ms'] / 1e3) / 1e6:.1f} M reads/s (kernels)", flush=True)
a, p, st = idx.coverage()
print("stats", st)
