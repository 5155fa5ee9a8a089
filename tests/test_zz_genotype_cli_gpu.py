"""GPU suite, sorted after every other file on purpose (written when the round's GPU minutes were spent: what it runs on
the device is what tests/test_gram_cli.py already runs; the genotyping half is host code, tested on the CPU in
tests/test_level_genotyper.py with these very inputs).

`gram genotype` end to end (genotype.cpp:24-118): quasimap on the GPU, then the genotyping step — the files under
geno_dir/genotype must be the ones the genotyper gives, run here on the CPU, for the ORACLE's coverage of the same reads."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, Oracle
from gramtools_b200 import encode_reads, level_genotype_json, master_seeds, read_depth_stats_host, synth

GRAM = os.path.join(ROOT, "gramtools_b200", "bin", "gram")


def _close(a, b):
    if isinstance(a, list):
        return len(a) == len(b) and all(_close(x, y) for x, y in zip(a, b))
    if isinstance(a, float) or isinstance(b, float):
        return abs(a - b) <= 1e-6 * max(1.0, abs(b))
    return a == b


def _run_case(tmp_path, name, prg, reads, k, seed, ploidy, debug=False):
    gd, od = tmp_path / f"{name}_gram", tmp_path / f"{name}_geno"
    gd.mkdir()
    np.asarray(prg, dtype="<u4").tofile(gd / "prg")
    fq = tmp_path / f"{name}.fq"
    fq.write_text("".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    cmd = [GRAM, "genotype", "--gram_dir", str(gd), "--reads", str(fq), "--sample_id", "smp", "--ploidy", ploidy,
           "--kmer_size", str(k), "--genotype_dir", str(od), "--max_threads", "2", "--seed", str(seed)]
    out = subprocess.run(cmd + (["--debug"] if debug else []), capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Running genotyping model" in out.stdout
    got = json.loads((od / "genotype" / "genotyped.json").read_text())

    draws = master_seeds(seed, 5000 * ((len(reads) + 4999) // 5000))
    bases, offs = encode_reads(reads)
    o = Oracle(prg, k)
    o.map(bases, offs, draws[:len(reads)], threads=4, want_states=False)
    res = o.result(want_states=False)
    depth = read_depth_stats_host(prg, res.per_base, res.grouped)
    rs = json.load(open(od / "read_stats.json"))
    assert abs(rs["Read_depth"]["Mean"] - depth["mean"]) <= 1e-5 * max(1.0, depth["mean"])
    assert rs["Read_depth"]["num_sites_total"] == depth["num_sites_total"]
    want = json.loads(level_genotype_json(prg, res.per_base, res.grouped, depth["mean"], depth["variance"], 1e-4,
                                          ploidy=ploidy, sample_id="smp"))
    for key in ("Child_Map", "Lvl1_Sites", "Model", "Samples", "Site_Fields", "Filters"):
        assert got[key] == want[key], key
    assert len(got["Sites"]) == len(want["Sites"])
    for i, (g, w) in enumerate(zip(got["Sites"], want["Sites"])):
        for key in ("ALS", "GT", "HAPG", "DP", "FT", "POS", "SEG"):
            assert g[key] == w[key], (name, i, key, g, w)
        for key in ("COV", "GT_CONF", "GT_CONF_PERCENTILE"):
            assert _close(g[key], w[key]), (name, i, key, g, w)
    vcf = gzip.decompress((od / "genotype" / "genotyped.vcf.gz").read_bytes()).decode().splitlines()
    n_lvl1 = len(got["Sites"]) if got["Lvl1_Sites"] == ["all"] else len(got["Lvl1_Sites"])
    assert len([l for l in vcf if not l.startswith("#")]) == n_lvl1
    assert (od / "genotype" / "personalised_reference.fasta").read_text().startswith(">gramtools_prg")
    if debug:
        assert (od / "site_gtyping_debug_info.txt").read_text().startswith("Model params:")
    return got


@pytest.mark.gpu
def test_gram_genotype_end_to_end_snp_prg(built_lib, tmp_path):
    prg, ref, pos, alt = synth.make_snp_prg(3000, 150, 9)
    haps = synth.snp_haplotypes(ref, pos, alt, 1, 10)
    b, o = synth.sample_reads(haps, 4000, 60, 11)
    reads = ["".join("?ACGT"[x] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    got = _run_case(tmp_path, "snp", prg, reads, 6, 7, "haploid", debug=True)
    assert sum(s["GT"] != [[None]] for s in got["Sites"]) >= 140  # one haplotype at 80x: the sites are called


@pytest.mark.gpu
def test_gram_genotype_end_to_end_nested_prg_diploid(built_lib, tmp_path):
    prg = synth.make_nested_prg(5, 300, 9)
    rng = np.random.default_rng(9)
    haps = [synth.random_haplotype(prg, rng) for _ in range(2)]
    b, o = synth.sample_reads(haps, 6000, 50, 10)
    reads = ["".join("?ACGT"[x] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    _run_case(tmp_path, "nested", prg, reads, 4, 3, "diploid")
