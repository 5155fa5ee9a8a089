"""GPU suite, last file of it: the k-mer index as gram_dir's sdsl files through the product (libgq.so, `gram build`,
`gram genotype --kmer_index_from_gram_dir`). The format itself is tested on the CPU in tests/test_sdsl_io.py."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, Oracle, assert_parity, gpu_result
from gramtools_b200 import QuasimapIndex, master_seeds, synth

GRAM = os.path.join(ROOT, "gramtools_b200", "bin", "gram")


def _reads_for(prg, n, L, seed):
    rng = np.random.default_rng(seed)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    return synth.sample_reads(haps, n, L, seed, frac_garbage=0.05, frac_n=0.02)


@pytest.mark.gpu
def test_kmer_index_files_round_trip_gpu(built_lib, tmp_path):
    """gq_kmer_index_dump -> gq_index_build_from_gram_dir (the k-mer index as gram_dir's sdsl files, kmer_index::load):
    the index loaded from the files maps like the oracle."""
    prg, k = synth.make_nested_prg(5, 300, 9), 4
    built = QuasimapIndex(prg, k, device=0)
    built.kmer_index_dump(str(tmp_path))
    for name in ("kmers", "kmers_stats", "sa_intervals", "paths"):
        assert os.path.getsize(tmp_path / name) > 8
    loaded = QuasimapIndex(prg, k, device=0, kmer_index_dir=str(tmp_path))
    assert loaded.layout.n_kmer_states == built.layout.n_kmer_states
    bases, offs = _reads_for(prg, 300, 40, 6)
    seeds = master_seeds(42, offs.size - 1)
    loaded.map_batch(bases, offs, seeds)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    assert_parity(gpu_result(loaded), o.result(), "index loaded from gram_dir files")
    built.close()
    loaded.close()


@pytest.mark.gpu
def test_gram_build_then_genotype_from_its_kmer_index(built_lib, tmp_path):
    """`gram build` leaves kmers / kmers_stats / sa_intervals / paths in gram_dir; `gram genotype
    --kmer_index_from_gram_dir` loads them (kmer_index::load, genotype.cpp:40) and writes the same coverage as a run
    that searches the k-mers again."""
    from gramtools_b200 import synth
    gram_dir = tmp_path / "gram"
    gram_dir.mkdir()
    prg = synth.make_snp_prg(3000, 100, 4)[0]
    np.asarray(prg, dtype="<u4").tofile(gram_dir / "prg")
    out = subprocess.run([GRAM, "build", "--gram_dir", str(gram_dir), "--kmer_size", "5"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("kmers", "kmers_stats", "sa_intervals", "paths"):
        assert (gram_dir / name).stat().st_size > 8
    rng = np.random.default_rng(1)
    hap = synth.random_haplotype(prg, rng)
    fq = tmp_path / "r.fastq"
    with open(fq, "w") as f:
        for i in range(200):
            s0 = int(rng.integers(0, hap.size - 60))
            f.write(f"@r{i}\n" + "".join("?ACGT"[x] for x in hap[s0:s0 + 60]) + "\n+\n" + "I" * 60 + "\n")
    dumps = []
    assert (gram_dir / "gq_index").stat().st_size > 1000
    for leg, extra in enumerate(([], ["--kmer_index_from_gram_dir"], ["--gq_index"])):
        geno = tmp_path / f"geno{leg}"
        out = subprocess.run([GRAM, "genotype", "--gram_dir", str(gram_dir), "--reads", str(fq), "--sample_id", "s",
                              "--ploidy", "haploid", "--kmer_size", "5", "--genotype_dir", str(geno), "--seed", "42"] + extra,
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        dumps.append([open(geno / "coverage" / n).read() for n in
                      ("allele_sum_coverage", "allele_base_coverage.json", "grouped_allele_counts_coverage.json")])
    assert dumps[0] == dumps[1] == dumps[2]
    # a stored index of another kmer_size is refused
    out = subprocess.run([GRAM, "genotype", "--gram_dir", str(gram_dir), "--reads", str(fq), "--sample_id", "s", "--ploidy",
                          "haploid", "--kmer_size", "6", "--genotype_dir", str(tmp_path / "g6"), "--gq_index"],
                         capture_output=True, text=True)
    assert out.returncode != 0 and "gram build" in out.stdout


@pytest.mark.gpu
def test_whole_index_file_round_trip_gpu(built_lib, tmp_path):
    """gq_index_save -> gq_index_load: the loaded index maps like the oracle and knows its PRG."""
    prg, k = synth.make_nested_prg(4, 300, 12), 4
    built = QuasimapIndex(prg, k, device=0)
    path = str(tmp_path / "gq_index")
    built.save(path)
    loaded = QuasimapIndex(None, 0, device=0, index_file=path)
    assert np.array_equal(loaded.prg(), np.asarray(prg, dtype=np.uint32))
    assert loaded.layout.kmer_size == k and loaded.layout.n_kmer_states == built.layout.n_kmer_states
    bases, offs = _reads_for(prg, 300, 40, 8)
    seeds = master_seeds(42, offs.size - 1)
    loaded.map_batch(bases, offs, seeds)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    assert_parity(gpu_result(loaded), o.result(), "index loaded from a gq_index file")
    built.close()
    loaded.close()
