"""The genotyping step behind quasimap (gramtools_b200/csrc/level_genotyper.cpp, host code of libgq.so).

* the reference's own unit expectations (tests/genotype/infer/**), transcribed in tests/genotyper/test_level_genotyper.cpp,
  run as a plain C++ program linked against level_genotyper.cpp;
* the reference's high-level cases (tests/genotype/infer/level_genotyping/test_runner.cpp:14-151): PRG -> quasimap (the
  oracle, CPU) -> read depth statistics -> LevelGenotyper through libgq.so's C entry points -> genotyped.json;
* the files of geno_dir/genotype: JSON layout, personalised reference, BGZF VCF.
No GPU: the reference runs this step on the host as well."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, Oracle
from gramtools_b200 import encode_reads, level_genotype, level_genotype_json, read_depth_stats_host


def test_reference_unit_expectations(built_lib):
    out = os.path.join(ROOT, "tests", "_build", "test_level_genotyper")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-Wall", "-fopenmp", "-o", out,
                    os.path.join(ROOT, "tests", "genotyper", "test_level_genotyper.cpp"),
                    os.path.join(ROOT, "gramtools_b200", "csrc", "level_genotyper.cpp"), "-lz"], check=True)
    r = subprocess.run([out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout and int(r.stdout.split()[0]) >= 150, r.stdout


def bracketed(s):
    """"AT[GC[C,A]T,TTA]T" -> PRG integers (sites 5, 7, … in opening order)"""
    out, opened, nxt = [], [], 5
    for c in s:
        if c == "[":
            opened.append(nxt)
            out.append(nxt)
            nxt += 2
        elif c == ",":
            out.append(opened[-1] + 1)
        elif c == "]":
            out.append(opened.pop() + 1)
        else:
            out.append("ACGT".index(c) + 1)
    return np.asarray(out, dtype=np.uint32)


def numbered(s):
    """"AA5C6G6AA" -> PRG integers (single-digit markers, as the reference's prg_string_to_ints takes them)"""
    return np.asarray([int(c) if c.isdigit() else "ACGT".index(c) + 1 for c in s], dtype=np.uint32)


def quasimap(prg, reads, k=2, forward_only=True):
    """forward_only: the reference's test helper maps the reads as given (prg_setup::quasimap_reads calls quasimap_read,
    test_resources.cpp:48-56); gram genotype maps both strands."""
    o = Oracle(prg, k)
    bases, offs = encode_reads(reads)
    if forward_only:
        o.map_forward(bases, offs, np.zeros(len(reads), dtype=np.uint32))
    else:
        o.map(bases, offs, np.zeros(len(reads), dtype=np.uint32), want_states=False)
    return o.result(want_states=False)


def genotype(prg, reads, qual_char, ploidy="haploid", k=2):
    res = quasimap(prg, reads, k)
    depth = read_depth_stats_host(prg, res.per_base, res.grouped)
    err = 10 ** (-(ord(qual_char) - 33) / 10) if reads else 0.0
    text = level_genotype_json(prg, res.per_base, res.grouped, depth["mean"], depth["variance"], err, ploidy=ploidy,
                               sample_id="s1")
    return json.loads(text), res, depth, text


def called(site):
    gt = site["GT"][0]
    return None if gt == [None] else [site["ALS"][g] for g in gt]


def test_two_site_non_nested_prg(built_lib):
    """LevelGenotyping.Given2SiteNonNestedPRG_CorrectGenotypes (test_runner.cpp:14-39)"""
    prg = numbered("AATAA5C6G6AA7C8G8AA")
    j, res, depth, text = genotype(prg, ["AATAACAACAA"] * 5 + ["AATAAGAACAA"], "?")
    assert [called(s) for s in j["Sites"]] == [["C"], ["C"]]
    assert [s["HAPG"] for s in j["Sites"]] == [[[0]], [[0]]]
    assert j["Sites"][0]["DP"] == [6] and j["Sites"][0]["COV"] == [[5.0]]
    assert j["Sites"][1]["DP"] == [6] and j["Sites"][1]["COV"] == [[6.0]]
    assert [s["POS"] for s in j["Sites"]] == [6, 9] and j["Sites"][0]["SEG"] == "gramtools_prg"
    assert j["Lvl1_Sites"] == ["all"] and j["Child_Map"] == {} and j["Model"] == "LevelGenotyping"
    assert j["Samples"] == [{"Desc": "made by gramtools genotype", "Name": "s1"}]
    assert depth["num_sites_total"] == 2 and depth["num_sites_noCov"] == 0
    assert depth["mean"] == 5.5 and depth["variance"] == 0.25
    # nlohmann::json streams objects with sorted keys and no white space (make_json.cpp, genotype.cpp:92-98)
    assert text.startswith('{"Child_Map":{},"Filters":{"AMBIG":{"Desc":"Ambiguous site.')
    assert '"Sites":[{"ALS":["C"],"COV":[[5.0]],"DP":[6],"FT":[[]],"GT":[[0]],"GT_CONF":[' in text
    assert sorted(j["Site_Fields"]) == ["ALS", "COV", "DP", "FT", "GT", "GT_CONF", "GT_CONF_PERCENTILE", "HAPG", "POS", "SEG"]
    for s in j["Sites"]:
        assert s["GT_CONF"][0] > 0 and 0 < s["GT_CONF_PERCENTILE"][0] <= 100
    # the confidence, recomputed independently (model.cpp:238-282): Poisson(5.5) depth model, error rate 1e-3,
    # log likelihood = wrong reads * ln(error) + ln pmf(mean allele coverage) (+ gap penalty: none here)
    from scipy.stats import poisson
    ll = lambda own, other: other * np.log(1e-3) + poisson.logpmf(own, 5.5)
    assert abs(j["Sites"][0]["GT_CONF"][0] - (ll(5, 1) - ll(1, 5))) < 1e-9
    assert abs(j["Sites"][1]["GT_CONF"][0] - (ll(6, 0) - (ll(0, 6) + poisson.logpmf(0, 5.5)))) < 1e-9  # G: 1 of 1 bases uncovered


def test_two_site_nested_prg(built_lib):
    """LevelGenotyping.Given2SiteNestedPRG_CorrectGenotypes (test_runner.cpp:41-65)"""
    prg = bracketed("AATAA[CCC[A,G],T]AA")
    j, res, _, _ = genotype(prg, ["AATAACCCGAA"] * 5 + ["AATAATAA"], "?")
    assert called(j["Sites"][1]) == ["G"] and j["Sites"][1]["HAPG"] == [[1]]
    assert called(j["Sites"][0]) == ["CCCG"] and j["Sites"][0]["HAPG"] == [[0]]
    # the REF allele CCCA was not called but is reported first (model.cpp:443-450)
    assert j["Sites"][0]["ALS"] == ["CCCA", "CCCG"] and j["Sites"][0]["GT"] == [[1]]
    assert j["Lvl1_Sites"] == [0] and j["Child_Map"] == {"0": {"0": [1]}}
    assert [s["POS"] for s in j["Sites"]] == [6, 9]


def test_direct_deletion_is_called(built_lib):
    """LevelGenotyper.GivenPRGWithDirectDeletion_CorrectlyCalledEmptyAllele (test_runner.cpp:67-88)"""
    prg = bracketed("GGGGG[CCC,]GG")
    j, _, _, _ = genotype(prg, ["GGGGGG"] * 5, "?")
    assert called(j["Sites"][0]) == [""] and j["Sites"][0]["HAPG"] == [[1]]
    assert j["Sites"][0]["ALS"] == ["CCC", ""]


SNPS_IN_TWO_HAPLOTYPES = "ATCGGC[TC[A,G]TC,GG[T,G]GG]AT"


def test_no_reads_all_null(built_lib):
    """LG_SnpsNestedInTwoHaplotypes.MapNoReads_AllGenotypesAreNull (test_runner.cpp:118-127)"""
    prg = bracketed(SNPS_IN_TWO_HAPLOTYPES)
    j, _, depth, _ = genotype(prg, [], ".")
    assert all(s["GT"] == [[None]] for s in j["Sites"])
    assert all(s["GT_CONF"] == [0.0] and s["DP"] == [0] and s["COV"] == [[]] for s in j["Sites"])
    assert [s["ALS"] for s in j["Sites"]] == [["TCATC"], ["A"], ["T"]]
    assert depth["num_sites_noCov"] == 1 and depth["num_sites_total"] == 1


def test_nested_sites_genotyped_and_invalidated(built_lib):
    """LG_SnpsNestedInTwoHaplotypes.MapReads_CorrectlyGenotypedSites / _CorrectlyInvalidatedSites (test_runner.cpp:129-151)"""
    prg = bracketed(SNPS_IN_TWO_HAPLOTYPES)
    j, _, _, _ = genotype(prg, ["ATCGGCTCGTCAT"] * 7 + ["ATCGGCGGG"], ".")
    assert called(j["Sites"][0]) == ["TCGTC"] and j["Sites"][0]["HAPG"] == [[0]]
    assert called(j["Sites"][1]) == ["G"] and j["Sites"][1]["HAPG"] == [[1]]
    # haplogroup 1 of site 0 was not called: the site living on it is nulled
    assert j["Sites"][2]["GT"] == [[None]] and j["Sites"][2]["GT_CONF"] == [0.0]
    assert j["Child_Map"] == {"0": {"0": [1], "1": [2]}} and j["Lvl1_Sites"] == [0]


def test_diploid_heterozygous_call(built_lib):
    prg = numbered("AATAA5C6G6AATT")
    j, _, _, _ = genotype(prg, ["AATAACAATT"] * 10 + ["AATAAGAATT"] * 9, "?", ploidy="diploid")
    s = j["Sites"][0]
    assert s["GT"] == [[0, 1]] and s["ALS"] == ["C", "G"] and s["HAPG"] == [[0, 1]] and s["COV"] == [[10.0, 9.0]]
    # one site: depth mean 10 (the C allele), variance 0 -> Poisson(10); het 0/1 against the best homozygous call
    # (model.cpp:284-335: a homozygous call halves the haplogroup's coverage between its two copies but scores each
    # copy on the allele's full mean per-base coverage)
    from scipy.stats import poisson
    lp = lambda c: poisson.logpmf(c, 10.0)
    het = 0 * np.log(1e-3) + lp(10) + lp(9)
    hom_c = 9 * np.log(1e-3) + 2 * lp(10)
    assert abs(s["GT_CONF"][0] - (het - hom_c)) < 1e-9


def test_ambiguous_site_is_filtered_and_filter_propagates(built_lib):
    """Two paths of the outer site spell the same sequence (model.cpp:30-31, runner.cpp:90-93)."""
    prg = bracketed("AAT[CC[A,G],CCA]TTA")
    j, _, _, _ = genotype(prg, ["AATCCGTTA"] * 6, "?")
    assert j["Sites"][0]["FT"] == [["AMBIG"]]   # duplicated candidate allele CCA
    assert j["Sites"][1]["FT"] == [["AMBIG"]]   # pushed down to the nested site
    assert called(j["Sites"][1]) == ["G"]


def test_output_files(built_lib, tmp_path):
    """genotype.cpp:85-118: genotyped.json, personalised_reference.fasta (distinct sequences, 60 columns),
    genotyped.vcf.gz (BGZF: any gzip reader takes it; level-1 sites only, 1-based positions per segment)."""
    prg = bracketed("AATAA[CCC[A,G],T]AA" + "ACGT" * 20 + "[C,G]TT")
    reads = ["AATAACCCGAAACGTACGTACG"] * 5 + ["GTACGTGTT"] * 5
    res = quasimap(prg, reads)
    depth = read_depth_stats_host(prg, res.per_base, res.grouped)
    coords = tmp_path / "prg_coords.tsv"
    coords.write_text("chr1\t20\nchr2\t76\n")
    out = tmp_path / "genotype"
    out.mkdir()
    debug = tmp_path / "site_gtyping_debug_info.txt"
    level_genotype(prg, res.per_base, res.grouped, depth["mean"], depth["variance"], 0.001, str(out), ploidy="diploid",
                   sample_id="smp", prg_coords_path=str(coords), debug_path=str(debug))
    j = json.loads((out / "genotyped.json").read_text())
    assert [(s["SEG"], s["POS"]) for s in j["Sites"]] == [("chr1", 6), ("chr1", 9), ("chr2", 72)]
    assert [s["GT"] for s in j["Sites"]] == [[[1, 1]], [[1, 1]], [[1, 1]]]
    fasta = (out / "personalised_reference.fasta").read_text()
    recs = fasta.split(">")[1:]
    seqs = {r.split("\n", 1)[0].split()[0]: r.split("\n", 1)[1].replace("\n", "") for r in recs}
    # both haplotypes are equal: one record per segment survives the deduplication
    assert len(recs) == 2 and all("smp personalised reference made by gramtools genotype" in r for r in recs)
    ref = "AATAA" + "CCCG" + "AA" + "ACGT" * 20 + "G" + "TT"
    assert sorted(seqs.values(), key=len) == sorted([ref[:20], ref[20:]], key=len)
    assert all(len(line) <= 60 for r in recs for line in r.split("\n")[1:])
    raw = (out / "genotyped.vcf.gz").read_bytes()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"  # BGZF member
    assert raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))  # BGZF EOF marker
    vcf = gzip.decompress(raw).decode().splitlines()
    assert vcf[0] == "##fileformat=VCFv4.2"
    assert '##contig=<ID=chr1,length=20,Source="gramtools">' in vcf and "##Model=LevelGenotyping" in vcf
    header = [l for l in vcf if l.startswith("#CHROM")]
    assert header == ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsmp"]
    body = [l.split("\t") for l in vcf if not l.startswith("#")]
    assert [(b[0], b[1], b[3], b[4]) for b in body] == [("chr1", "6", "CCCA", "CCCG"), ("chr2", "72", "C", "G")]
    assert all(b[8] == "GT:DP:COV:FT:GT_CONF:GT_CONF_PERCENTILE" for b in body)
    # the uncalled REF shares haplogroup 0 with the called allele: it is given half of that haplogroup's own coverage
    assert body[0][9].startswith("1/1:5:2.5,5:PASS:") and body[1][9].startswith("1/1:5:0,5:PASS:")
    text = debug.read_text()
    assert text.startswith("Model params: \nmean cov: ") and "site index: \t0" in text


def test_depth_statistics_follow_the_most_covered_path(built_lib):
    """read_stats.cpp:72-160 on a nested PRG: the walk takes, in every site it meets, the haplogroup with the highest
    grouped count; a site whose best path holds no base contributes that count."""
    prg = bracketed("ATCGGC[TC[A,G]TC,GG[T,G]GG]AT" + "CCCC[TTT,]GG")
    reads = ["ATCGGCTCGTCAT"] * 7 + ["ATCGGCGGG"] + ["CCCCGG"] * 3
    res = quasimap(prg, reads)
    d = read_depth_stats_host(prg, res.per_base, res.grouped)
    # site 0: path TC G TC with 7 reads on every base -> 7; site 3: direct deletion with 3 reads -> 3
    assert d["num_sites_total"] == 2 and d["num_sites_noCov"] == 0
    assert d["mean"] == 5.0 and d["variance"] == 4.0


def test_result_does_not_depend_on_threads(built_lib, tmp_path):
    """Level-1 sites (with what is nested in them) are genotyped in parallel: same files for any thread count."""
    from gramtools_b200 import master_seeds, synth
    for prg, k, seed in ((synth.make_snp_prg(6000, 400, 2)[0], 6, 2), (synth.make_nested_prg(90, 120, 4), 5, 4)):
        rng = np.random.default_rng(seed)
        haps = [synth.random_haplotype(prg, rng) for _ in range(2)]
        b, o = synth.sample_reads(haps, 8000, 50, seed + 1)
        orc = Oracle(prg, k)
        orc.map(b, o, master_seeds(seed, 8000), threads=4, want_states=False)
        res = orc.result(want_states=False)
        depth = read_depth_stats_host(prg, res.per_base, res.grouped)
        files = []
        for nt in (1, 4, 0):
            out = tmp_path / f"g{seed}_{nt}"
            out.mkdir()
            level_genotype(prg, res.per_base, res.grouped, depth["mean"], depth["variance"], 1e-3, str(out),
                           ploidy="diploid", n_threads=nt)
            files.append([(out / f).read_bytes() for f in ("genotyped.json", "personalised_reference.fasta", "genotyped.vcf.gz")])
        assert files[0] == files[1] == files[2]
        assert len(json.loads(files[0][0])["Sites"]) >= 90


def test_reference_read_mapping_stats(built_lib):
    """ReadMappingStats.GivenThreeMappedReadsNonNestedPRG_… / GivenTwoMappedReadsNestedPRG_… (test_read_stats.cpp:140-183):
    quasimap (oracle) -> ReadStats::compute_coverage_depth."""
    prg = numbered("G5CAAA6AA6T7G8C8GGG")
    res = quasimap(prg, ["AAA", "AAA", "GCAAA", "GCAAA"])
    d = read_depth_stats_host(prg, res.per_base, res.grouped)
    assert (d["mean"], d["variance"], d["num_sites_noCov"], d["num_sites_total"]) == (1.75, 3.0625, 1, 2)
    prg = bracketed("G[GG[G,A]G,C]CCC")
    res = quasimap(prg, ["GGGGGCCC", "GCCCC", "GCCCC", "GCCC"])
    d = read_depth_stats_host(prg, res.per_base, res.grouped)
    assert (d["mean"], d["variance"], d["num_sites_noCov"], d["num_sites_total"]) == (3.0, 0.0, 0, 1)


def test_bad_input_is_refused(built_lib):
    from gramtools_b200 import GqError
    prg = numbered("AATAA5C6G6AA")
    with pytest.raises(GqError, match="per-base vector"):
        level_genotype_json(prg, np.zeros(5, np.uint16), np.zeros(0, np.uint32), 10, 0, 0.01)
    with pytest.raises(GqError, match="bad record|truncated"):
        level_genotype_json(prg, np.zeros(2, np.uint16), np.asarray([7, 1, 1, 0], np.uint32), 10, 0, 0.01)
    with pytest.raises(GqError):
        level_genotype_json(numbered("AA5C6G"), np.zeros(2, np.uint16), np.zeros(0, np.uint32), 10, 0, 0.01)


def _cli_like(prg, reads, k, seeds, tmp_path, name, ploidy="haploid"):
    """What `gram genotype` hands to the genotyper for these reads (qualities 'I' = Q40): oracle coverage in place of
    the GPU's (bit-identical by the parity tests), depth statistics, error rate 1e-4."""
    bases, offs = encode_reads(reads)
    o = Oracle(prg, k)
    o.map(bases, offs, np.asarray(seeds, dtype=np.uint32), threads=4, want_states=False)
    res = o.result(want_states=False)
    depth = read_depth_stats_host(prg, res.per_base, res.grouped)
    out = tmp_path / name
    out.mkdir()
    level_genotype(prg, res.per_base, res.grouped, depth["mean"], depth["variance"], 1e-4, str(out), ploidy=ploidy,
                   sample_id="s")
    j = json.loads((out / "genotyped.json").read_text())
    vcf = gzip.decompress((out / "genotyped.vcf.gz").read_bytes()).decode().splitlines()
    fasta = (out / "personalised_reference.fasta").read_text()
    return j, vcf, fasta, res, depth


def test_inputs_of_the_cli_tests(built_lib, tmp_path):
    """The PRGs and reads the GPU tests push through `gram genotype` (tests/test_gram_cli.py,
    tests/test_kmer_index_files_gpu.py), genotyped here on the CPU from the oracle's coverage: the step runs to the end
    on every one of them, calls what the reads were drawn from, and its three files agree with one another."""
    from gramtools_b200 import master_seeds, synth
    dec = lambda b, o: ["".join("?ACGT"[x] for x in b[int(o[i]):int(o[i + 1])]) for i in range(o.size - 1)]
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json")))
    for name, case in fx.items():
        prg = np.asarray(case["prg"], dtype=np.uint32)
        for ploidy in ("haploid", "diploid"):
            j, vcf, fasta, res, _ = _cli_like(prg, case["reads"], case["kmer_size"], master_seeds(42, len(case["reads"])),
                                              tmp_path, f"{name}_{ploidy}", ploidy)
            assert len(j["Sites"]) >= 1 and fasta.startswith(">gramtools_prg")
            lvl1 = len(j["Sites"]) if j["Lvl1_Sites"] == ["all"] else len(j["Lvl1_Sites"])
            assert len([l for l in vcf if not l.startswith("#")]) == lvl1

    # SNP PRG, reads from four haplotypes (test_cli_matches_oracle_two_files_gz_and_seed_batches)
    prg, ref, pos, alt = synth.make_snp_prg(3000, 150, 9)
    haps = synth.snp_haplotypes(ref, pos, alt, 4, 10)
    b1, o1 = synth.sample_reads(haps, 7000, 60, 11)
    b2, o2 = synth.sample_reads(haps, 3000, 60, 12)
    r1, r2 = dec(b1, o1), dec(b2, o2)
    r2[5] = r2[5][:10] + "N" + r2[5][11:]
    draws = master_seeds(7, 20000)
    j, vcf, fasta, res, depth = _cli_like(prg, r1 + r2, 6, np.concatenate([draws[:7000], draws[10000:13000]]), tmp_path, "snp")
    assert len(j["Sites"]) == 150 and j["Lvl1_Sites"] == ["all"] and depth["num_sites_total"] == 150
    body = [l.split("\t") for l in vcf if not l.startswith("#")]
    assert len(body) == 150 and [int(b[1]) for b in body] == sorted(int(b[1]) for b in body)
    assert [b[1] for b in body] == [str(s["POS"]) for s in j["Sites"]]
    n_called = sum(s["GT"] != [[None]] for s in j["Sites"])
    assert n_called >= 100  # four haplotypes in a haploid model: sites where they disagree may stay uncalled
    for s, b in zip(j["Sites"], body):
        assert b[3] == s["ALS"][0] and (b[4] == "." if len(s["ALS"]) == 1 else b[4] == ",".join(s["ALS"][1:]))
    pers = "".join(fasta.split("\n")[1:])
    assert len(pers) == len(ref)  # SNPs only: the personalised reference keeps the length

    # one haplotype, 200 reads (test_gram_build_then_genotype_from_its_kmer_index): every covered site is called as drawn
    prg = synth.make_snp_prg(3000, 100, 4)[0]
    rng = np.random.default_rng(1)
    hap = synth.random_haplotype(prg, rng)
    reads = []
    for i in range(200):
        s0 = int(rng.integers(0, hap.size - 60))
        reads.append("".join("?ACGT"[x] for x in hap[s0:s0 + 60]))
    j, vcf, fasta, res, depth = _cli_like(prg, reads, 5, master_seeds(42, 200), tmp_path, "one_hap")
    pers = "".join(fasta.split("\n")[1:])
    hap_s = "".join("?ACGT"[x] for x in hap)
    assert len(pers) == len(hap_s)
    called_sites = [s for s in j["Sites"] if s["GT"] != [[None]]]
    assert len(called_sites) >= 80
    for s in called_sites:  # 1-based POS of a SNP site: the personalised and the true haplotype agree there
        assert pers[s["POS"] - 1] == hap_s[s["POS"] - 1] == s["ALS"][s["GT"][0][0]]

    # nested PRG (test_cli_two_devices_matches_one, test_kmer_index_files_round_trip_gpu shapes)
    for n_loci, seed, k in ((6, 17, 5), (5, 9, 4), (4, 12, 4)):
        prg = synth.make_nested_prg(n_loci, 300, seed)
        rng = np.random.default_rng(seed)
        haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
        b, o = synth.sample_reads(haps, 6000, 50, seed + 1)
        for ploidy in ("haploid", "diploid"):
            j, vcf, fasta, res, depth = _cli_like(prg, dec(b, o), k, master_seeds(3, 6000), tmp_path,
                                                  f"nested{seed}_{ploidy}", ploidy)
            assert j["Lvl1_Sites"] != ["all"] and len(j["Child_Map"]) >= 1
            assert len([l for l in vcf if not l.startswith("#")]) == len(j["Lvl1_Sites"])
            # a site nested on a haplogroup that its parent did not call is null (runner.cpp:135-187)
            for parent, by_hapg in j["Child_Map"].items():
                ps = j["Sites"][int(parent)]
                if ps["GT"] == [[None]]:
                    continue
                for hapg, kids in by_hapg.items():
                    if int(hapg) not in ps["HAPG"][0]:
                        assert all(j["Sites"][c]["GT"] == [[None]] for c in kids)


def test_prg_without_sites(built_lib, tmp_path):
    """No site: depth statistics are 0 / 0 (as the reference computes them), nothing to genotype, the personalised
    reference is the PRG itself and the VCF has only its header."""
    prg = np.asarray([1, 2, 3, 4, 1, 1], dtype=np.uint32)
    none16, none32 = np.zeros(0, np.uint16), np.zeros(0, np.uint32)
    d = read_depth_stats_host(prg, none16, none32)
    assert np.isnan(d["mean"]) and d["num_sites_total"] == 0
    j = json.loads(level_genotype_json(prg, none16, none32, d["mean"], d["variance"], 0.001))
    assert j["Sites"] == [] and j["Lvl1_Sites"] == ["all"]
    level_genotype(prg, none16, none32, d["mean"], d["variance"], 0.001, str(tmp_path))
    assert (tmp_path / "personalised_reference.fasta").read_text().split("\n")[1] == "ACGTAA"
    assert gzip.decompress((tmp_path / "genotyped.vcf.gz").read_bytes()).decode().splitlines()[-1].startswith("#CHROM")
