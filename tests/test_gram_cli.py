"""The process seam: `gram genotype` argv contract + geno_dir layout (genotype.py:71-93,
parameters.cpp:45-116), served by libgq.so."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, Oracle
from gramtools_b200 import encode_reads, master_seeds, synth

GRAM = os.path.join(ROOT, "gramtools_b200", "bin", "gram")


def _write_prg(path, prg):
    np.asarray(prg, dtype="<u4").tofile(path)


def _fastq(path, reads, gz=False):
    txt = "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads))
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(txt)
    else:
        open(path, "w").write(txt)


def _run(gram_dir, geno_dir, reads, k, seed=42, extra=(), threads=1, env=None):
    cmd = [GRAM, "genotype", "--gram_dir", str(gram_dir), "--reads", *[str(r) for r in reads], "--sample_id", "s",
           "--ploidy", "haploid", "--kmer_size", str(k), "--genotype_dir", str(geno_dir), "--max_threads", str(threads),
           "--seed", str(seed), *extra]
    return subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))


def _grouped(json_path):
    d = json.load(open(json_path))["grouped_allele_counts"]
    return [{",".join(map(str, d["allele_groups"][g])): c for g, c in site.items()} for site in d["site_counts"]]


def test_gram_without_arguments_exits_zero(built_lib):
    # gramtools_main.py:80-90 runs the bare executable to check it works
    out = subprocess.run([GRAM], capture_output=True, text=True)
    assert out.returncode == 0 and "genotype" in out.stdout


def test_gram_genotype_missing_arguments_exits_nonzero(built_lib):
    out = subprocess.run([GRAM, "genotype", "--gram_dir", "x"], capture_output=True, text=True)
    assert out.returncode == 1 and "required" in out.stdout
    out = subprocess.run([GRAM, "frobnicate"], capture_output=True, text=True)
    assert out.returncode == 1


def test_gram_build_arguments_and_no_cpu_fallback(built_lib, tmp_path):
    out = subprocess.run([GRAM, "build"], capture_output=True, text=True)
    assert out.returncode == 1 and "--kmer_size" in out.stdout
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    np.asarray([1, 2, 5, 3, 6, 4, 6, 1, 2, 3, 4, 1], dtype="<u4").tofile(tmp_path / "prg")
    out = subprocess.run([GRAM, "build", "--gram_dir", str(tmp_path), "--kmer_size", "3", "--max_threads", "2"],
                         capture_output=True, text=True)
    assert out.returncode != 0 and "no CUDA device" in (out.stdout + out.stderr)
    assert "Loaded PRG: 12 symbols" in out.stdout  # the whole file in one read, a trailing partial word ignored
    with open(tmp_path / "prg", "ab") as f:
        f.write(b"\x01\x00")
    out = subprocess.run([GRAM, "build", "--gram_dir", str(tmp_path), "--kmer_size", "3"], capture_output=True, text=True)
    assert "Loaded PRG: 12 symbols" in out.stdout
    assert not (tmp_path / "kmers").exists()


@pytest.mark.gpu
def test_integration_fixtures_through_cli(built_lib, tmp_path):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json")))
    for name, case in fx.items():
        gd, od = tmp_path / f"{name}_gram", tmp_path / f"{name}_geno"
        gd.mkdir()
        _write_prg(gd / "prg", case["prg"])
        _fastq(tmp_path / f"{name}.fq", case["reads"])
        out = _run(gd, od, [tmp_path / f"{name}.fq"], case["kmer_size"])
        assert out.returncode == 0, out.stdout + out.stderr
        pb = json.load(open(od / "coverage" / "allele_base_coverage.json"))["allele_base_counts"]
        assert pb == case.get("allele_base_counts", [])
        got = _grouped(od / "coverage" / "grouped_allele_counts_coverage.json")
        exp = [case["grouped"].get(str(s), {}) for s in range(len(got))]
        assert got == exp, name
        rs = json.load(open(od / "read_stats.json"))
        assert rs["Quality"]["Num_bases"] == sum(len(r) for r in case["reads"])
        assert abs(rs["Quality"]["Error_rate_mean"] - 10 ** (-4.0)) < 1e-9   # all qualities 'I' = Q40
        assert "Count exact mapped reads" in out.stdout


@pytest.mark.gpu
def test_cli_matches_oracle_two_files_gz_and_seed_batches(built_lib, tmp_path):
    """Two read files (one gzipped, one with N reads): read j of a file gets the j-th draw of its
    5000-read buffer and unused draws of a partly filled buffer are discarded (quasimap.cpp:132-139)."""
    prg, ref, pos, alt = synth.make_snp_prg(3000, 150, 9)
    haps = synth.snp_haplotypes(ref, pos, alt, 4, 10)
    b1, o1 = synth.sample_reads(haps, 7000, 60, 11)
    b2, o2 = synth.sample_reads(haps, 3000, 60, 12)
    dec = lambda b, o: ["".join("?ACGT"[x] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    r1, r2 = dec(b1, o1), dec(b2, o2)
    r2[5] = r2[5][:10] + "N" + r2[5][11:]
    gd, od = tmp_path / "gram", tmp_path / "geno"
    gd.mkdir()
    _write_prg(gd / "prg", prg)
    _fastq(tmp_path / "a.fq.gz", r1, gz=True)
    _fastq(tmp_path / "b.fq", r2)
    out = _run(gd, od, [tmp_path / "a.fq.gz", tmp_path / "b.fq"], 6, seed=7)
    assert out.returncode == 0, out.stdout + out.stderr
    draws = master_seeds(7, 20000)
    seeds = np.concatenate([draws[:7000], draws[10000:13000]])   # file 1 consumed 2 buffers = 10000 draws
    bases, offs = encode_reads(r1 + r2)
    o = Oracle(prg, 6)
    o.map(bases, offs, seeds, threads=4, want_states=False)
    ref_r = o.result(want_states=False)
    lines = [l.split() for l in open(od / "coverage" / "allele_sum_coverage").read().strip().split("\n")]
    assert [int(x) for l in lines for x in l] == list(ref_r.allele_sum)
    pb = json.load(open(od / "coverage" / "allele_base_coverage.json"))["allele_base_counts"]
    assert [c for s in pb for a in s for c in a] == list(ref_r.per_base)
    got = _grouped(od / "coverage" / "grouped_allele_counts_coverage.json")
    exp = [dict() for _ in got]
    w, i = [int(x) for x in ref_r.grouped], 0
    while i < len(w):
        n = w[i + 2]
        exp[w[i]][",".join(map(str, w[i + 3:i + 3 + n]))] = w[i + 1]
        i += 3 + n
    assert got == exp
    assert f"Count all reads: {ref_r.stats[0]}" in out.stdout
    assert f"Count skipped reads with no sequence: {ref_r.stats[1]}" in out.stdout
    assert f"Count exact mapped reads: {ref_r.stats[4]}" in out.stdout


@pytest.mark.gpu
def test_cli_two_devices_matches_one(built_lib, tmp_path):
    """`--devices 2`: reads packed by host threads and sharded over two GPUs batch by batch, one NCCL exchange at the
    end — every output file identical to the single-GPU run (and so to the oracle, checked above)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    prg = synth.make_nested_prg(6, 300, 17)
    rng = np.random.default_rng(17)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    b, o = synth.sample_reads(haps, 30000, 50, 18)
    reads = ["".join("?ACGT"[x] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    gd = tmp_path / "gram"
    gd.mkdir()
    _write_prg(gd / "prg", prg)
    _fastq(tmp_path / "r.fq", reads)
    outs = []
    for nd in (1, 2):
        od = tmp_path / f"geno{nd}"
        r = _run(gd, od, [tmp_path / "r.fq"], 5, seed=3, extra=["--devices", str(nd)], threads=4,
                 env={"GQ_BATCH_READS": "4096"})  # several batches, so both GPUs get work
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append({f: open(od / "coverage" / f).read() for f in
                     ("allele_sum_coverage", "allele_base_coverage.json", "grouped_allele_counts_coverage.json")})
        outs[-1]["counts"] = [ln for ln in r.stdout.splitlines() if ln.startswith("Count ")]
    assert outs[0]["allele_sum_coverage"] == outs[1]["allele_sum_coverage"]
    assert outs[0]["counts"] == outs[1]["counts"]
    assert _grouped(tmp_path / "geno1" / "coverage" / "grouped_allele_counts_coverage.json") == \
        _grouped(tmp_path / "geno2" / "coverage" / "grouped_allele_counts_coverage.json")
