"""The `gram` executable's HOST side, end to end, on a machine without a GPU.

`gram genotype` / `gram build` are the product's process seam; their device half is libgq.so's kernels (GPU tests), their
host half — argument handling, FASTQ / gz reader, 5000-draw seed rule, batch queue, packer hand-over, coverage dumps,
read statistics, the genotyping step, the index files — is ordinary C++ that the GPU suite also runs, but only on the
GPU box. Here the same executable runs under LD_PRELOAD of tests/emu/gq_shim.cpp: the device-bound entry points of
include/gq.h are answered by the host emulation of the device functions (test infrastructure, the code
tests/test_host_parity.py holds against the oracle), everything else by the real libgq.so. The bodies of the GPU CLI
tests are run unchanged. The product itself has no CPU path (tests/test_capi_symbols.py, test_gram_cli.py)."""
import json
import os
import subprocess

import pytest

import test_gram_cli as cli
import test_kmer_index_files_gpu as files_gpu
import test_zz_genotype_cli_gpu as genotype_gpu
from common import ROOT


@pytest.fixture(scope="module")
def shim(built_lib):
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu"), "../_build/libgq_shim.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "tests", "_build", "libgq_shim.so")


@pytest.fixture
def preloaded(shim, monkeypatch):
    monkeypatch.setenv("LD_PRELOAD", shim)
    return shim


def test_integration_fixtures(built_lib, tmp_path, preloaded):
    cli.test_integration_fixtures_through_cli(built_lib, tmp_path)
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json")))
    for name in fx:  # and the genotyping step ran to the end on each of them
        j = json.load(open(tmp_path / f"{name}_geno" / "genotype" / "genotyped.json"))
        assert j["Model"] == "LevelGenotyping" and len(j["Sites"]) >= 1
        assert (tmp_path / f"{name}_geno" / "genotype" / "genotyped.vcf.gz").stat().st_size > 28
        assert (tmp_path / f"{name}_geno" / "genotype" / "personalised_reference.fasta").read_text().startswith(">")


def test_two_files_gz_and_seed_batches(built_lib, tmp_path, preloaded):
    cli.test_cli_matches_oracle_two_files_gz_and_seed_batches(built_lib, tmp_path)


def test_build_then_genotype_from_its_files(built_lib, tmp_path, preloaded):
    files_gpu.test_gram_build_then_genotype_from_its_kmer_index(built_lib, tmp_path)


def test_genotype_end_to_end(built_lib, tmp_path, preloaded):
    genotype_gpu.test_gram_genotype_end_to_end_snp_prg(built_lib, tmp_path)
    genotype_gpu.test_gram_genotype_end_to_end_nested_prg_diploid(built_lib, tmp_path)


def test_small_batches_and_threads(built_lib, tmp_path, preloaded, monkeypatch):
    """Many small batches through the queue (GQ_BATCH_READS), several packing threads: same files as one batch."""
    import numpy as np
    from gramtools_b200 import synth
    prg = synth.make_nested_prg(4, 300, 12)
    rng = np.random.default_rng(12)
    haps = [synth.random_haplotype(prg, rng) for _ in range(3)]
    b, o = synth.sample_reads(haps, 3000, 50, 13, frac_garbage=0.05, frac_n=0.02)
    reads = ["".join("?ACGTN"[min(int(x), 5)] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    gd = tmp_path / "gram"
    gd.mkdir()
    cli._write_prg(gd / "prg", prg)
    cli._fastq(tmp_path / "r.fq", reads)
    outs = []
    for leg, (batch, threads) in enumerate((("1048576", 1), ("257", 3))):
        od = tmp_path / f"geno{leg}"
        r = cli._run(gd, od, [tmp_path / "r.fq"], 4, seed=5, threads=threads, env={"GQ_BATCH_READS": batch})
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append({f: open(od / sub / f).read() for sub, f in
                     (("coverage", "allele_sum_coverage"), ("coverage", "grouped_allele_counts_coverage.json"),
                      ("genotype", "genotyped.json"), ("genotype", "personalised_reference.fasta"))})
        outs[-1]["counts"] = [ln for ln in r.stdout.splitlines() if ln.startswith("Count ")]
    assert outs[0] == outs[1]


def test_argv_of_the_python_front_end(built_lib, tmp_path, preloaded):
    """The exact command lines gramtools' Python front-end builds (commands/build/build.py:65-80,
    commands/genotype/genotype.py:71-93), --debug included: build leaves the k-mer index files, genotype the whole
    geno_dir layout (paths.py:131-152) that genotype.py reads next."""
    import numpy as np
    from gramtools_b200 import synth
    gd, od = tmp_path / "gram", tmp_path / "geno"
    gd.mkdir()
    prg = synth.make_snp_prg(2000, 60, 3)[0]
    cli._write_prg(gd / "prg", prg)
    (gd / "prg_coords.tsv").write_text("ref1\t2000\n")
    rng = np.random.default_rng(3)
    hap = synth.random_haplotype(prg, rng)
    reads = ["".join("?ACGT"[x] for x in hap[s0:s0 + 70]) for s0 in rng.integers(0, hap.size - 70, 600)]
    cli._fastq(tmp_path / "r.fq", reads)
    out = subprocess.run([cli.GRAM, "build", "--gram_dir", str(gd), "--ref", str(tmp_path / "ref.fa"), "--kmer_size", "5",
                          "--max_threads", "2", "--all_kmers", "--debug"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert all((gd / f).stat().st_size > 8 for f in ("kmers", "kmers_stats", "sa_intervals", "paths"))
    out = subprocess.run([cli.GRAM, "genotype", "--gram_dir", str(gd), "--reads", str(tmp_path / "r.fq"), "--sample_id",
                          "sample one", "--ploidy", "diploid", "--kmer_size", "5", "--genotype_dir", str(od),
                          "--max_threads", "2", "--seed", "42", "--debug"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for f in ("read_stats.json", "site_gtyping_debug_info.txt", "coverage/allele_sum_coverage",
              "coverage/allele_base_coverage.json", "coverage/grouped_allele_counts_coverage.json",
              "genotype/genotyped.json", "genotype/genotyped.vcf.gz", "genotype/personalised_reference.fasta"):
        assert (od / f).stat().st_size > 0, f
    rs = json.load(open(od / "read_stats.json"))
    assert rs["Read_depth"]["num_sites_total"] == 60 and rs["Read_depth"]["num_sites_noCov"] < 30  # genotype.py:95-118
    j = json.load(open(od / "genotype" / "genotyped.json"))
    assert j["Samples"][0]["Name"] == "sample one" and all(s["SEG"] == "ref1" for s in j["Sites"])
    fasta = (od / "genotype" / "personalised_reference.fasta").read_text()
    assert fasta.startswith(">ref1_1 sample one personalised reference made by gramtools genotype\n")
    assert len("".join(fasta.split(">")[1].split("\n")[1:])) == 2000
