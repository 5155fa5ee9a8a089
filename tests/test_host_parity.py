"""CPU suite: the product's host code (flat index builder, k-mer index) and — through the test-only
host emulation of the device functions (tests/emu) — the per-strand search/coverage logic, against
the oracle. Bit-exact: integer path, no tolerance."""
import json
import os

import ctypes as C

import numpy as np
import pytest

from common import ROOT, Emu, Oracle, assert_parity, reference_test_cases, uint16_case
from gramtools_b200 import encode_reads, master_seeds, synth


def _reads_for(prg, n, L, seed, garbage=0.05, n_frac=0.02):
    rng = np.random.default_rng(seed)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    return synth.sample_reads(haps, n, L, seed, frac_garbage=garbage, frac_n=n_frac)


def _check(prg, k, bases, offs, seed=42, arena_words=256, what=""):
    seeds = master_seeds(seed, offs.size - 1)
    o, e = Oracle(prg, k), Emu(prg, k)
    assert (o.n_sites, o.n_alleles, o.n_per_base, o.is_nested, o.sa_size, o.n_kmer_states) == \
           (e.n_sites, e.n_alleles, e.n_per_base, e.is_nested, e.sa_size, e.n_kmer_states)
    o.map(bases, offs, seeds)
    e.map(bases, offs, seeds, arena_words=arena_words)
    ro, re = o.result(), e.result()
    assert_parity(re, ro, what)
    return ro, re


@pytest.mark.parametrize("seed", range(4))
def test_snp_prg_parity(seed):
    prg, _, _, _ = synth.make_snp_prg(800, 60, seed)
    bases, offs = _reads_for(prg, 500, 50, seed)
    ro, _ = _check(prg, 5, bases, offs, what=f"snp{seed}")
    assert ro.stats[4] > 100 and ro.allele_sum.sum() > 0 and ro.per_base.sum() > 0


@pytest.mark.parametrize("seed", range(6))
def test_nested_prg_parity(seed):
    prg = synth.make_nested_prg(3, 250, seed)
    bases, offs = _reads_for(prg, 500, 30, seed)
    ro, _ = _check(prg, 4, bases, offs, what=f"nested{seed}")
    assert ro.stats[4] > 100 and ro.grouped.size > 0


@pytest.mark.parametrize("seed", range(3))
def test_indel_prg_parity(seed):
    """config 4/5 site mix: SNPs + deletions + insertions (ambiguous placements -> multi-allele groups)."""
    prg = synth.make_indel_prg(1500, 80, seed)
    bases, offs = _reads_for(prg, 500, 45, seed)
    ro, _ = _check(prg, 5, bases, offs, what=f"indel{seed}")
    assert ro.stats[4] > 100


def test_tiny_arena_forces_overflow_reruns():
    """Arena overflow must re-run the strand with a larger arena, never truncate."""
    prg = synth.make_nested_prg(2, 200, 11)
    bases, offs = _reads_for(prg, 300, 25, 11, garbage=0, n_frac=0)
    _, re = _check(prg, 3, bases, offs, arena_words=40, what="tiny-arena")
    assert re.extra["reruns"] > 0


def test_edge_cases():
    prg = np.asarray(synth.make_snp_prg(300, 20, 3)[0])
    k = 5
    rng = np.random.default_rng(0)
    hap = synth.random_haplotype(prg, rng)
    s = lambda a: "".join("?ACGT"[x] for x in a)
    reads = [
        "",                       # empty read -> skipped
        "ACGTNACGTACGT",          # non-ACGT -> emptied -> skipped (utils.cpp:72-81)
        "ACG",                    # shorter than k
        s(hap[:k]),               # exactly k
        s(hap[10:10 + k + 1]),
        s(hap[:120]),             # long read
        s(hap[-40:]),             # touches the PRG end
        s(hap[:40]),              # touches the PRG start
        "A" * 30, "acgtacgtacgtacgt",
    ]
    bases, offs = encode_reads(reads)
    ro, _ = _check(prg, k, bases, offs, what="edges")
    assert ro.stats[0] == 2 * len(reads) and ro.stats[1] == 4
    # ragged batch with zero reads
    bases, offs = encode_reads([])
    _check(prg, k, bases, offs, what="empty-batch")


def test_sparse_marker_ids():
    """Alphabet compression (sdsl's char2comp) with marker ids that are far apart — the presence-table branch and, for
    large ids, the sort-based one: the suffix array is the brute-force one and the flat index passes its invariants.
    (Coverage for such PRGs is undefined in the reference: its vectors are indexed by (site - 5) / 2.)"""
    base = np.asarray(synth.make_snp_prg(400, 12, 8)[0]).astype(np.int64)
    for stride, offset in ((1, 0), (3, 0), (50, 1000), (20000, 0)):
        prg = base.copy()
        m = prg >= 5
        site = (prg[m] - 5) // 2
        prg[m] = 5 + 2 * (site * stride + offset) + ((prg[m] - 5) % 2)
        prg = prg.astype(np.uint32)
        e = Emu(prg, 4)
        e.index_check()
        text = np.concatenate([prg.astype(np.int64), [0]])
        want = sorted(range(text.size), key=lambda i: text[i:].tolist())
        assert e.sa().tolist() == want, (stride, offset)


def test_reference_test_prgs():
    """PRGs + reads of the reference's quasimap tests (test_quasimap.cpp), both strands, incl. the seed-dependent
    selections (seeds 42 / 150 / 29 / 200)."""
    for name, prg, k, reads, seeds in reference_test_cases():
        bases, offs = encode_reads(reads)
        for seed in seeds:
            _check(prg, k, bases, offs, seed=seed, what=name)


def test_uint16_wrap_and_saturation():
    """allele_sum and grouped counts wrap mod 65536, per-base counts saturate at 65535 — the oracle counts in
    uint16_t like the reference; the product counts in uint32 and converts when fetching."""
    prg, k, reads = uint16_case()
    bases, offs = encode_reads(reads)
    ro, re = _check(prg, k, bases, offs, what="uint16")
    assert ro.per_base.max() == 65535 and (ro.per_base == 65535).sum() >= 3       # saturated cells
    assert 0 < ro.allele_sum[1] < 10000 and ro.stats[4] >= 136000                  # 70000 wrapped to 4464
    g = [int(x) for x in ro.grouped]
    recs, i = {}, 0
    while i < len(g):
        recs[(g[i], tuple(g[i + 3:i + 3 + g[i + 2]]))] = g[i + 1]
        i += 3 + g[i + 2]
    assert recs[(1, (0, 1))] == 66000 - 65536 and recs[(0, (1,))] == 70000 - 65536


def test_group_table_growth():
    """A multi-allele group table that is far too small must grow (flagged strands are given back uncommitted
    and recorded after the rebuild), never drop or double count."""
    prg = synth.make_nested_prg(6, 300, 4)
    bases, offs = _reads_for(prg, 1500, 30, 4, garbage=0.0, n_frac=0.0)
    seeds = master_seeds(42, offs.size - 1)
    o, e = Oracle(prg, 4), Emu(prg, 4)
    e.set_gtab_cap(4)
    o.map(bases, offs, seeds)
    e.map(bases, offs, seeds)
    ro, re = o.result(), e.result()
    assert_parity(re, ro, "gtab-growth")
    n_multi, g, i = 0, [int(x) for x in ro.grouped], 0
    while i < len(g):
        n_multi += g[i + 2] > 1
        i += 3 + g[i + 2]
    assert n_multi > 16 and re.extra["reruns"] > 0, (n_multi, re.extra)


def test_integration_fixtures_through_product_code():
    with open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json")) as f:
        fx = json.load(f)
    for name, case in fx.items():
        bases, offs = encode_reads(case["reads"])
        _check(np.asarray(case["prg"], dtype=np.uint32), case["kmer_size"], bases, offs, what=name)


def test_suffix_array_matches_oracle_order():
    prg = synth.make_nested_prg(4, 300, 5)
    e = Emu(prg, 3)
    sa = e.sa()
    text = np.concatenate([prg, [0]]).astype(np.int64)
    assert sorted(sa.tolist()) == list(range(text.size))
    # adjacent suffixes strictly increasing
    for a, b in zip(sa[:-1], sa[1:]):
        ta, tb = text[a:], text[b:]
        m = min(ta.size, tb.size)
        d = np.nonzero(ta[:m] != tb[:m])[0]
        assert d.size and ta[d[0]] < tb[d[0]]


def test_suffix_array_64bit_indices(monkeypatch):
    """SA-IS with 64-bit indices (the instantiation texts beyond 2^31 symbols take, BASELINE config 5) gives the
    suffix array of the 32-bit one — on random texts over small and large alphabets, on runs, and on PRGs; and an
    index built through it maps like the oracle."""
    from common import emu_lib
    lib = emu_lib()
    rng = np.random.default_rng(5)
    texts = []
    for sigma, n in ((2, 1), (3, 2), (3, 500), (5, 4000), (50, 3000), (1000, 20000), (6, 100000)):
        t = rng.integers(1, sigma, size=n).astype(np.int32) if sigma > 2 else np.ones(n, dtype=np.int32)
        texts.append((np.concatenate([t, [0]]).astype(np.int32), sigma))
    texts.append((np.concatenate([np.tile([1, 2], 5000), [0]]).astype(np.int32), 3))  # deep recursion: periodic text
    texts.append((np.concatenate([np.full(10000, 3), [0]]).astype(np.int32), 4))
    for prg in (synth.make_nested_prg(11, 400, 6), synth.make_snp_prg(5000, 200, 3)[0]):
        present = np.unique(prg)
        comp = np.searchsorted(present, prg).astype(np.int32) + 1
        texts.append((np.concatenate([comp, [0]]).astype(np.int32), int(present.size) + 1))
    for t, sigma in texts:
        t = np.ascontiguousarray(t)
        assert lib.emu_sais64_agrees(t.ctypes.data_as(C.POINTER(C.c_int32)), t.size, sigma) == 1, (t.size, sigma)
    monkeypatch.setenv("GQ_SAIS64", "1")
    prg = synth.make_nested_prg(2, 300, 4)
    bases, offs = _reads_for(prg, 200, 40, 9)
    _check(prg, 4, bases, offs, what="index built with 64-bit SA-IS")


def test_malformed_prgs_rejected():
    for bad in ([5, 1, 6, 2, 6, 5, 3, 6, 4, 6], [1, 5, 2, 6, 4], [1, 5, 6, 2], [1, 6, 2, 6], [1, 5, 2, 6, 3]):
        with pytest.raises(RuntimeError):
            Emu(np.asarray(bad, dtype=np.uint32), 2)


def test_text_and_general_routes_agree():
    """The text-mode fast path (seed candidates -> verify -> text walk) and the general lane machine must
    give the same states and coverage; both are exercised, and both match the oracle."""
    for name, prg, k, L in (("snp", synth.make_snp_prg(3000, 200, 5)[0], 6, 70),
                            ("nested", synth.make_nested_prg(6, 300, 5), 5, 40),
                            ("indel", synth.make_indel_prg(3000, 150, 5), 6, 60)):
        bases, offs = _reads_for(prg, 1500, L, 9, garbage=0.02, n_frac=0.0)
        seeds = master_seeds(42, offs.size - 1)
        o, e = Oracle(prg, k), Emu(prg, k)
        o.map(bases, offs, seeds)
        ro = o.result()
        e.routes(reset=True)
        e.map(bases, offs, seeds)
        r = e.routes()
        assert_parity(e.result(), ro, name + "/default")
        assert r["fast_finished"] > 0, r
        if name == "snp":
            assert r["fast_finished"] > 20 * r["general"], r  # isolated SNPs: nearly everything on the fast path
        if name == "nested":
            assert r["general"] > 0, r  # adjacent / nested markers need the general jump machinery
        g = Emu(prg, k)
        g.force_general(True)
        try:
            g.map(bases, offs, seeds)
            r2 = g.routes()
        finally:
            g.force_general(False)
        assert r2["fast_finished"] == 0 and r2["general"] > 0, r2
        assert_parity(g.result(), ro, name + "/general-only")


def test_repeats_take_the_general_route():
    """A read that maps to two copies of a repeat is ONE SearchState with a 2-wide interval in the reference;
    the split candidates both finish, so the strand must be redone by the general machinery."""
    rng = np.random.default_rng(3)
    unit = rng.integers(1, 5, 120)
    prg = np.concatenate([rng.integers(1, 5, 200), unit, rng.integers(1, 5, 150), unit, rng.integers(1, 5, 200)]).astype(np.uint32)
    s = lambda a: "".join("?ACGT"[x] for x in a)
    reads = [s(unit[10:70]), s(unit[30:110]), s(prg[150:230]), s(prg[5:90])]
    bases, offs = encode_reads(reads)
    seeds = master_seeds(1, len(reads))
    o, e = Oracle(prg, 6), Emu(prg, 6)
    o.map(bases, offs, seeds)
    e.routes(reset=True)
    e.map(bases, offs, seeds)
    r = e.routes()
    assert_parity(e.result(), o.result(), "repeats")
    assert r["multi_finisher"] >= 2 and r["general"] >= 2, r


@pytest.mark.parametrize("shape", [(200000, 400, 5, 60), (400000, 4000, 6, 70), (300000, 300, 4, 50)])
def test_frequent_kmers_wide_seed_intervals(shape):
    """Large-genome regime in miniature (config 4/5: tens of occurrences per k-mer): the seeding k-mer's main state
    is an interval of ~100-200 suffixes, enumerated in the seed view of the index; in the last shape (~1200
    occurrences per k-mer, beyond kSplitWidth) the seed pass narrows it with rank steps, peeling off
    marker-preceded suffixes as candidates of their own."""
    n, n_sites, k, L = shape
    prg = synth.make_snp_prg(n, n_sites, 7)[0]
    bases, offs = _reads_for(prg, 1500, L, 5, garbage=0.02, n_frac=0.0)
    seeds = master_seeds(42, offs.size - 1)
    o, e = Oracle(prg, k), Emu(prg, k)
    o.map(bases, offs, seeds, threads=os.cpu_count())
    e.routes(reset=True)
    e.map(bases, offs, seeds)
    r = e.routes()
    assert_parity(e.result(), o.result(), f"wide{shape}")
    assert r["fast_finished"] > 2 * r["general"] and r["too_wide"] == 0, r


def test_large_k_seed_buckets():
    """k = 12: the seed view is bucketed by ONE context base (k <= 11: two, k >= 13: none) — index invariants and
    parity on an indel PRG. (k = 13 was run by hand when the buckets were written: the 4^13 tables make it a
    30-second oracle build.)"""
    prg = synth.make_indel_prg(3000, 150, 5)
    bases, offs = _reads_for(prg, 400, 60, 3, garbage=0.05, n_frac=0.01)
    Emu(prg, 12).index_check()
    ro, _ = _check(prg, 12, bases, offs, what="k12")
    assert ro.stats[4] > 100


def test_flat_index_invariants():
    """Text groups, inverse SA, text-order jump records, the seed view (per-suffix entries with left context),
    the reverse-complement presence set and the inline first edge: checked structurally on SNP, nested, indel and
    frequent-k-mer PRGs (k-mers with more than kSplitWidth occurrences keep an interval entry)."""
    for prg, k in ((synth.make_snp_prg(3000, 200, 5)[0], 6), (synth.make_nested_prg(6, 300, 5), 5),
                   (synth.make_indel_prg(3000, 150, 5), 6), (synth.make_snp_prg(60000, 100, 2)[0], 3),
                   (np.asarray([1, 2, 3, 4, 5, 1, 6, 2, 6, 3, 3, 7, 4, 8, 8, 1], dtype=np.uint32), 2)):
        Emu(prg, k).index_check()


def test_kmer_index_holds_the_oracles_states():
    """index_kmers (build.cpp:101-131): k-mer by k-mer, the product's index holds exactly the SearchStates the oracle's
    vBWT searches end with — SA interval, traversed loci in order, traversing sites — in the reference's list order:
    the extended input states first, the marker-derived ones after them (vBWT_jump.cpp:119-132), markers taken from
    the highest SA index down and an entered / exited state committed before the loci chained to it (the LIFO worklist,
    :155-183). quasimap never reads the order (classes are selected through a std::map, coverage_common.cpp:166-177);
    it is the order of the records in gram_dir's sa_intervals / paths files."""
    def records(words):
        out, i = [], 0
        while i < len(words):
            n = 5 + 2 * words[i + 3] + 2 * words[i + 4]
            out.append(tuple(words[i:i + n]))
            i += n
        return out
    cases = [(synth.make_snp_prg(2000, 80, 3)[0], 5), (synth.make_indel_prg(2000, 60, 5), 5),
             (synth.make_snp_prg(30000, 40, 8)[0], 3),  # wide intervals: many markers per state
             (synth.make_nested_prg(5, 300, 9), 4), (synth.make_nested_prg(8, 200, 3, max_depth=4), 3),
             (np.asarray([1, 2, 3, 4, 5, 1, 6, 2, 6, 3, 3, 7, 4, 8, 8, 1], dtype=np.uint32), 2)]
    cases += [(synth.make_nested_prg(6, 250, 20 + sd), 3 + sd % 3) for sd in range(6)]
    cases += [(np.asarray(c["prg"], dtype=np.uint32), c["kmer_size"]) for c in
              json.load(open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json"))).values()]
    cases += [(prg, k) for _, prg, k, _, _ in reference_test_cases()]  # test_quasimap.cpp: adjacent sites, direct
    for prg, k in cases:                                                # deletions, doubly nested sites, empty alleles
        want, got = records(Oracle(prg, k).kmer_states()), records(Emu(prg, k).kmer_states())
        assert len(want) > 0 and want == got


def test_long_reads_host():
    prg = synth.make_snp_prg(30000, 1200, 21)[0]
    bases, offs = _reads_for(prg, 300, 700, 21, garbage=0.3, n_frac=0.0)
    ro, _ = _check(prg, 8, bases, offs, what="long-reads")
    assert ro.stats[4] > 100


def test_random_shapes():
    """Seeded sweep over random PRG shapes (isolated SNPs, nested bracket grammar, indels, duplicated segments),
    k in 2..7, read lengths from k to 90 (so reads with L == k and reads shorter than the walk buffers both occur),
    1 % substitution errors in a third of the cases, arenas down to 64 words: emulation == oracle, and the flat
    index passes its structural checks. (3000 cases of this sweep were run when the text-mode route was written;
    60 are kept here.)"""
    for case in range(60):
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
        bases, offs = _reads_for(prg, 300, L, case, garbage=0.05, n_frac=0.01)
        if float(rng.choice([0.0, 0.0, 0.01])) > 0:
            bases = bases.copy()
            hit = rng.random(bases.size) < 0.01
            bases[hit] = rng.integers(1, 5, int(hit.sum()))
        seeds = master_seeds(int(rng.integers(0, 1000)), offs.size - 1)
        o, e = Oracle(prg, k), Emu(prg, k)
        e.index_check()
        o.map(bases, offs, seeds)
        e.map(bases, offs, seeds, arena_words=int(rng.choice([64, 256, 1024])))
        assert_parity(e.result(), o.result(), f"random-shape-{case}")


def test_index_independent_of_threads_and_sa_builder():
    """The flat index is a pure function of (PRG, k): the parallel passes of the builder (rank blocks, marker records,
    ISA, k-mer subtrees, seed view) give the same bytes on 1 and on 8 threads, and with the 64-bit SA-IS."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from common import Emu\nfrom gramtools_b200 import synth\n"
            "for prg, k in ((synth.make_nested_prg(12, 600, 5), 6), (synth.make_snp_prg(120000, 4000, 2)[0], 8),\n"
            "               (synth.make_indel_prg(30000, 1500, 3), 7)):\n"
            "    print(Emu(prg, k).index_digest())\n") % (ROOT, os.path.join(ROOT, "tests"))
    outs = []
    for env_extra in ({"OMP_NUM_THREADS": "1"}, {"OMP_NUM_THREADS": "8"}, {"OMP_NUM_THREADS": "3", "GQ_SAIS64": "1"}):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.split())
    assert len(outs[0]) == 3 and outs[0] == outs[1] == outs[2], outs
