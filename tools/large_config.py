"""Developer run of a large SNP configuration on one GPU (e.g. the config-4 shape: 250 Mb, 5 M sites,
k = 11): index build time, kernel times, reads/s, and size-independent sanity properties. Not part of
the default bench (the host index build takes minutes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gramtools_b200 import QuasimapIndex, synth  # noqa: E402

ref_len, n_sites, k, n_reads = [int(x) for x in sys.argv[1:5]]
t = time.time()
prg, ref, pos, alt = synth.make_snp_prg(ref_len, n_sites, 0x6772616D + 4)
print(f"PRG: {prg.size} symbols, {pos.size} sites ({time.time() - t:.1f} s)", flush=True)
t = time.time()
idx = QuasimapIndex(prg, k)
lay = idx.layout
print(f"index build+upload {time.time() - t:.1f} s; device bytes {lay.device_bytes / 1e9:.2f} GB; "
      f"k-mer states {lay.n_kmer_states}", flush=True)
haps = synth.snp_haplotypes(ref, pos, alt, 4, 11)
bases, offs = synth.sample_reads(haps, n_reads, 150, 12)
seeds = synth.master_seeds(42, n_reads)
idx.upload(bases, offs, seeds)
for it in range(4):
    idx.map_resident()
    info = idx.run_info()
    print(info, f"-> {n_reads / (info['kernels_ms'] / 1e3) / 1e6:.1f} M reads/s (kernels)", flush=True)
status = idx.batch_status().reshape(-1, 2)
a, p, st = idx.coverage()
print("stats", st)
assert ((status == 3).sum(axis=1) >= 1).all(), "an error-free read failed to map"
assert st.all_reads_count == 2 * n_reads * 4
print("allele_sum total", int(a.astype(np.int64).sum()), "per_base total", int(p.astype(np.int64).sum()))
