"""GPU suite (`-m gpu`): the CUDA path through the C ABI (libgq.so) against the oracle on the same
seeded inputs — final SearchStates (SA intervals + both paths), all three coverage structures and the
five counters, bit-exact."""
import json
import os

import numpy as np
import pytest

from common import ROOT, Oracle, assert_parity, gpu_result, reference_test_cases, uint16_case
from gramtools_b200 import QuasimapIndex, encode_reads, master_seeds, synth

pytestmark = pytest.mark.gpu


def _reads_for(prg, n, L, seed, garbage=0.05, n_frac=0.02):
    rng = np.random.default_rng(seed)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    return synth.sample_reads(haps, n, L, seed, frac_garbage=garbage, frac_n=n_frac)


def _check(prg, k, bases, offs, seed=42, what="", options=None, threads=1):
    seeds = master_seeds(seed, offs.size - 1)
    idx = QuasimapIndex(prg, k, device=0)
    for name, val in (options or {}).items():
        idx.set_option(name, val)
    idx.map_batch(bases, offs, seeds)
    got = gpu_result(idx)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds, threads=threads)
    ref = o.result()
    assert_parity(got, ref, what)
    idx.close()
    return got, ref


@pytest.mark.parametrize("seed", range(3))
def test_snp_prg(built_lib, seed):
    prg, _, _, _ = synth.make_snp_prg(2000, 120, seed)
    bases, offs = _reads_for(prg, 3000, 60, seed)
    got, ref = _check(prg, 5, bases, offs, what=f"snp{seed}")
    assert ref.stats[4] > 1000


@pytest.mark.parametrize("seed", range(4))
def test_nested_prg(built_lib, seed):
    prg = synth.make_nested_prg(4, 300, seed)
    bases, offs = _reads_for(prg, 3000, 30, seed)
    got, ref = _check(prg, 4, bases, offs, what=f"nested{seed}")
    assert ref.grouped.size > 0


@pytest.mark.parametrize("seed", range(2))
def test_indel_prg(built_lib, seed):
    prg = synth.make_indel_prg(6000, 300, seed)
    bases, offs = _reads_for(prg, 4000, 70, seed)
    _check(prg, 6, bases, offs, what=f"indel{seed}", threads=4)


def test_nested_prg_larger(built_lib):
    """config 3 shape, scaled to what the oracle checks in seconds: nested loci, many multi-state reads."""
    prg = synth.make_nested_prg(40, 400, 77)
    bases, offs = _reads_for(prg, 30000, 60, 77, garbage=0.02, n_frac=0.0)
    got, ref = _check(prg, 7, bases, offs, what="nested-large", threads=8)
    assert ref.stats[4] > 10000


def test_config1_toy(built_lib):
    """BASELINE config 1: 1 kb ref + 50 biallelic SNPs, 10k x 100 bp reads, k=5."""
    prg, ref, pos, alt = synth.make_snp_prg(1000, 50, 0x6772616D)
    haps = synth.snp_haplotypes(ref, pos, alt, 8, 1)
    bases, offs = synth.sample_reads(haps, 10000, 100, 2)
    got, _ = _check(prg, 5, bases, offs, what="config1", threads=8)
    assert got.stats[0] == 20000


def test_overflow_reruns(built_lib):
    prg = synth.make_nested_prg(3, 250, 11)
    bases, offs = _reads_for(prg, 2000, 25, 11, garbage=0, n_frac=0)
    got, _ = _check(prg, 3, bases, offs, what="tiny-arena", options={"arena_words": 40})
    assert got.extra["rerun_strands"] > 0


def test_edge_cases_and_accumulation(built_lib):
    prg = np.asarray(synth.make_snp_prg(300, 20, 3)[0])
    rng = np.random.default_rng(0)
    hap = synth.random_haplotype(prg, rng)
    s = lambda a: "".join("?ACGT"[x] for x in a)
    reads = ["", "ACGTNACGTACGT", "ACG", s(hap[:5]), s(hap[10:16]), s(hap[:120]), s(hap[-40:]), s(hap[:40]), "A" * 30]
    bases, offs = encode_reads(reads)
    _check(prg, 5, bases, offs, what="edges")
    # coverage accumulates across batches; empty batch is a no-op
    idx = QuasimapIndex(prg, 5)
    o = Oracle(prg, 5)
    for b in range(3):
        bb, oo = _reads_for(prg, 500, 40, 100 + b)
        sd = master_seeds(7 + b, 500)
        idx.map_batch(bb, oo, sd)
        o.map(bb, oo, sd)
    e_b, e_o = encode_reads([])
    idx.map_batch(e_b, e_o, np.zeros(0, np.uint32))
    got = gpu_result(idx)
    ref = o.result()
    assert got.stats == ref.stats
    assert np.array_equal(got.allele_sum, ref.allele_sum) and np.array_equal(got.per_base, ref.per_base)
    assert np.array_equal(got.grouped, ref.grouped)
    idx.reset_coverage()
    a, p, st = idx.coverage()
    assert a.sum() == 0 and p.sum() == 0 and st.all_reads_count == 0


def test_integration_fixtures(built_lib):
    with open(os.path.join(ROOT, "tests", "golden", "it_fixtures.json")) as f:
        fx = json.load(f)
    for name, case in fx.items():
        bases, offs = encode_reads(case["reads"])
        got, _ = _check(np.asarray(case["prg"], dtype=np.uint32), case["kmer_size"], bases, offs, what=name)
        if "allele_base_counts" in case:
            flat = [c for site in case["allele_base_counts"] for allele in site for c in allele]
            assert list(got.per_base) == flat
        # grouped counts against the golden expectation itself (test_genotype_integration_tests.py:68-157)
        w, i, groups = [int(x) for x in got.grouped], 0, {}
        while i < len(w):
            n = w[i + 2]
            groups.setdefault(str(w[i]), {})[",".join(map(str, w[i + 3:i + 3 + n]))] = w[i + 1]
            i += 3 + n
        assert groups == case["grouped"], name


def test_medium_snp_properties(built_lib):
    """Larger than the oracle comfortably checks state-by-state: coverage parity without states plus
    size-independent properties (counters add up; every error-free read maps on exactly one strand
    or both; allele_sum >= grouped singles)."""
    prg, ref, pos, alt = synth.make_snp_prg(200_000, 4000, 5)
    haps = synth.snp_haplotypes(ref, pos, alt, 8, 6)
    bases, offs = synth.sample_reads(haps, 100_000, 150, 7)
    seeds = master_seeds(42, 100_000)
    idx = QuasimapIndex(prg, 8)
    idx.map_batch(bases, offs, seeds)
    got = gpu_result(idx)
    st = got.stats
    assert st[0] == 200_000 and st[1] + st[2] + st[3] + st[4] == st[0]
    status = got.status.reshape(-1, 2)
    assert ((status == 3).sum(axis=1) >= 1).all(), "an error-free read failed to map on either strand"
    o = Oracle(prg, 8)
    o.map(bases, offs, seeds, threads=8, want_states=False)
    ref_r = o.result(want_states=False)
    assert_parity(got, ref_r, "medium", check_states=False)


def test_multi_gpu_plumbing_single_device(built_lib):
    """The pieces bench.py uses for N>1: device pointers of the dense accumulators (aliased by a torch
    tensor for the NCCL all-reduce) and export/import of sparse multi-allele groups."""
    import torch
    from gramtools_b200.distributed import finalize_counters, shard_bounds
    prg = synth.make_nested_prg(3, 250, 31)
    bases, offs = _reads_for(prg, 2000, 30, 31)
    seeds = master_seeds(42, 2000)
    parts = []
    for r in range(2):  # two "ranks" on one device
        lo, hi = shard_bounds(2000, r, 2)
        idx = QuasimapIndex(prg, 4)
        idx.map_batch(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], seeds[lo:hi])
        ptr, n_cnt, _ = idx.device_counters()

        class _Shim:
            __cuda_array_interface__ = {"shape": (n_cnt,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Shim(), device="cuda:0").clone()
        parts.append((t, idx.groups_export(), idx))
    total = (parts[0][0] + parts[1][0]).cpu().numpy().astype(np.uint32)
    merged = QuasimapIndex(prg, 4)
    merged.groups_import(parts[0][1], replace=True)
    merged.groups_import(parts[1][1], replace=False)
    groups = merged.groups_export()
    o = Oracle(prg, 4)
    o.map(bases, offs, seeds)
    ref = o.result()
    a, p, g = finalize_counters(total, o.n_alleles, o.n_per_base, groups, parts[0][2].allele_offsets())
    assert np.array_equal(a, ref.allele_sum) and np.array_equal(p, ref.per_base)
    assert np.array_equal(g, ref.grouped)


@pytest.mark.parametrize("options", [{"seed_pass": 0}, {"seed_recs_per_read": 1}, {"resident_slices": 3}])
def test_route_options(built_lib, options):
    """General kernel only / candidate pool too small (strands fall back to the general kernel) / slices on two
    streams: same results."""
    for prg, k, L in ((synth.make_snp_prg(3000, 200, 5)[0], 6, 70), (synth.make_nested_prg(6, 300, 5), 5, 40)):
        bases, offs = _reads_for(prg, 4000, L, 9, garbage=0.02, n_frac=0.0)
        _check(prg, k, bases, offs, what=str(options), options=options, threads=4)


def test_repeats(built_lib):
    rng = np.random.default_rng(3)
    unit = rng.integers(1, 5, 120)
    prg = np.concatenate([rng.integers(1, 5, 200), unit, rng.integers(1, 5, 150), unit, rng.integers(1, 5, 200)]).astype(np.uint32)
    s = lambda a: "".join("?ACGT"[x] for x in a)
    reads = [s(unit[10:70]), s(unit[30:110]), s(prg[150:230]), s(prg[5:90])] * 50
    bases, offs = encode_reads(reads)
    _check(prg, 6, bases, offs, what="repeats")


def test_config2_full_size_properties(built_lib):
    """BASELINE config 2 at full size (4.4 Mb + 100k SNPs, 1M x 150 bp, k=10): too big for the oracle, so
    size-independent properties: every error-free read maps exactly one strand (random 4.4 Mb reference: no
    150-mer repeats), each mapped strand adds one allele_sum count per site it crosses and one per-base count per
    allele base it covers (SNP alleles are single bases, so the two totals agree), a second pass doubles every
    counter (idempotent accumulation), the pipelined host path equals the resident path, and a 20k-read prefix
    is bit-identical to the oracle."""
    import bench
    n = bench.N_READS
    prg, bases, offs, seeds = bench.make_workload(0, n)
    idx = QuasimapIndex(prg, bench.KMER, device=0)
    idx.upload(bases, offs, seeds)
    idx.map_resident()
    status = idx.batch_status().reshape(-1, 2)
    assert ((status == 3).sum(axis=1) == 1).all()
    a1, p1, st1 = idx.coverage()
    assert st1.all_reads_count == 2 * n and st1.exact_mapped_reads_count == n and st1.skipped_reads_count == 0
    assert st1.missing_kmer_reads_count + st1.no_extension_reads_count == n
    tot = int(a1.astype(np.int64).sum())
    assert tot == int(p1.astype(np.int64).sum()) and tot > 3 * n
    g1 = idx.grouped()
    idx.map_resident()
    a2, p2, st2 = idx.coverage()
    assert np.array_equal(a2.astype(np.int64), 2 * a1.astype(np.int64)) and np.array_equal(p2.astype(np.int64), 2 * p1.astype(np.int64))
    assert st2.exact_mapped_reads_count == 2 * n
    idx.reset_coverage()
    idx.map_batch(bases, offs, seeds)  # host buffers, sliced + pipelined over two streams
    a3, p3, st3 = idx.coverage()
    assert np.array_equal(a3, a1) and np.array_equal(p3, p1) and st3 == st1
    assert np.array_equal(idx.grouped(), g1)
    idx.close()
    m = 20000
    _check(prg, bench.KMER, bases[:int(offs[m])], offs[:m + 1], what="config2-prefix", threads=os.cpu_count())


def test_frequent_kmers_wide_seed_intervals(built_lib):
    """config 4/5 regime in miniature: ~200 occurrences per k-mer, wide seed intervals narrowed in the seed pass."""
    prg = synth.make_snp_prg(200000, 400, 7)[0]
    bases, offs = _reads_for(prg, 6000, 60, 5, garbage=0.02, n_frac=0.0)
    _check(prg, 5, bases, offs, what="wide-seeds", threads=os.cpu_count())
    prg = synth.make_snp_prg(300000, 300, 7)[0]  # ~1200 occurrences per 4-mer: narrowed with rank steps
    bases, offs = _reads_for(prg, 3000, 50, 5, garbage=0.02, n_frac=0.0)
    _check(prg, 4, bases, offs, what="very-wide-seeds", threads=os.cpu_count())


def test_long_reads(built_lib):
    """Reads of more than 512 bases (more than 32 packed words): the classify kernel then takes its k-mer windows
    from memory instead of the lanes' registers; long walks, many sites per read."""
    prg = synth.make_snp_prg(30000, 1200, 21)[0]
    bases, offs = _reads_for(prg, 1500, 700, 21, garbage=0.3, n_frac=0.0)
    got, ref = _check(prg, 8, bases, offs, what="long-reads-k8", threads=os.cpu_count())
    assert ref.stats[2] > 0 and ref.stats[4] > 500  # unmappable strands miss a k-mer (8-mers are sparse here)
    got, ref = _check(prg, 6, bases, offs, what="long-reads-k6", threads=os.cpu_count())
    assert ref.stats[3] > 0 and ref.stats[4] > 500  # every 6-mer occurs: the whole strand is probed, no extension


def test_random_shapes_gpu(built_lib):
    """The seeded sweep of tests/test_host_parity.py::test_random_shapes through libgq (first 24 cases; case 7 —
    3-base reads, k = 2, a PRG with duplicated segments — has dozens of finished candidates per strand and once
    overran the per-strand coverage work list). tools/gpu_fuzz.py runs more cases."""
    for case in range(24):
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
        _check(prg, k, bases, offs, seed=int(rng.integers(0, 1000)), what=f"random-shape-{case}",
               options={"arena_words": int(rng.choice([64, 256, 1024]))})


def test_reference_seeded_selection_gpu(built_lib):
    """The reference's quasimap test PRGs through libgq.so, including the seed-dependent multi-class selections
    with seeds 42 / 150 / 29 / 200 (test_quasimap.cpp:174-198,240-258,386-404) and the nested bracket PRGs."""
    for name, prg, k, reads, seeds in reference_test_cases():
        bases, offs = encode_reads(reads)
        for seed in seeds:
            _check(prg, k, bases, offs, seed=seed, what=f"{name}/seed{seed}")


def test_uint16_wrap_and_saturation_gpu(built_lib):
    """Counters past 65535: allele_sum and grouped counts (single- and multi-allele) wrap, per-base saturates
    (allele_sum.cpp:40-41, grouped_allele_counts.cpp:44-47, allele_base.cpp:239-241); expected values from the
    oracle, which counts in uint16_t like the reference."""
    prg, k, reads = uint16_case()
    bases, offs = encode_reads(reads)
    got, ref = _check(prg, k, bases, offs, what="uint16", threads=os.cpu_count())
    assert got.per_base.max() == 65535 and (got.per_base == 65535).sum() >= 3
    assert got.allele_sum[1] == 70000 - 65536
    # the same batch mapped in two halves on the same handle accumulates to the same wrapped totals
    idx = QuasimapIndex(prg, k)
    seeds = master_seeds(42, offs.size - 1)
    h = (offs.size - 1) // 2
    idx.map_batch(bases[:int(offs[h])], offs[:h + 1], seeds[:h])
    idx.map_batch(bases[int(offs[h]):], offs[h:] - offs[h], seeds[h:])
    a, p, _ = idx.coverage()
    assert np.array_equal(a, ref.allele_sum) and np.array_equal(p, ref.per_base)
    assert np.array_equal(idx.grouped(), ref.grouped)


def test_group_table_growth_gpu(built_lib):
    """Multi-allele group table of 4 slots on a nested PRG with dozens of distinct groups: the table grows
    (strands flagged before anything is committed, re-run after the rebuild) and results stay exact."""
    prg = synth.make_nested_prg(6, 300, 4)
    bases, offs = _reads_for(prg, 6000, 30, 4, garbage=0.0, n_frac=0.0)
    got, _ = _check(prg, 4, bases, offs, what="gtab-growth", options={"gtab_cap": 4}, threads=4)
    assert got.extra["rerun_strands"] > 0


def test_state_pool_overflow_multislice(built_lib):
    """Final-state pool far too small on a repeat-rich nested PRG, mapped in several pipelined slices: text-kernel
    winners and general-kernel strands both overflow, some strands are flagged twice; every strand must still be
    mapped and recorded exactly once."""
    rng = np.random.default_rng(8)
    unit = synth.make_nested_prg(2, 200, 8)
    # two copies of the same flanks around distinct loci: reads in the flanks have several finished candidates
    flank = rng.integers(1, 5, 300).astype(np.uint32)
    renum = unit.copy()
    renum[renum > 4] += int(unit.max()) - 4  # site ids continue after the first copy's (max marker is even)
    prg = np.concatenate([flank, unit, flank, renum, flank]).astype(np.uint32)
    bases, offs = _reads_for(prg, 40000, 40, 8, garbage=0.01, n_frac=0.0)
    got, _ = _check(prg, 5, bases, offs, what="pool-overflow",
                    options={"pool_words_per_read": 1, "chunk_reads": 4096, "tail_chunk_reads": 1024}, threads=os.cpu_count())
    assert got.extra["rerun_strands"] > 1000


@pytest.mark.parametrize("distinct", [True, False])
def test_config3_shape_prefix(built_lib, distinct):
    """BASELINE config 3 at its stated shape (nested PRG, 200 loci x 5 kb, k = 10, 150 bp reads): a 20k-read
    prefix bit-identical to the oracle — states, all three coverage structures, counters. distinct=True is the
    bench's PRG (pairwise distinct alleles, as make_prg builds them); distinct=False lets short alleles coincide,
    which multiplies the search states of the reads crossing such sites (up to hundreds per strand): the worst
    case for the general kernel, the general coverage route and their overflow re-runs."""
    prg = synth.make_nested_prg(200, 5000, 0x6772616D + 3, distinct=distinct)
    rng = np.random.default_rng(3)
    haps = [synth.random_haplotype(prg, rng) for _ in range(8)]
    bases, offs = synth.sample_reads(haps, 20000, 150, 13)
    got, ref = _check(prg, 10, bases, offs, what="config3-shape", threads=os.cpu_count())
    assert ref.stats[4] >= 20000 and ref.grouped.size > 0
    if not distinct:
        assert ref.state_count.max() >= 64


def test_config4_shape_prefix(built_lib):
    """BASELINE config 4's regime scaled so that the oracle builds its index in seconds: the same site mix (80 %
    SNPs, 10 % deletions, 10 % insertions, one site per 50 bp) and the same ~60 occurrences per seeding k-mer
    (config 4: 2.7e8 symbols / 4^11; here 4 Mb / 4^8), 150 bp reads: 20k-read prefix bit-identical to the oracle."""
    prg, ref, sites = synth.make_indel_prg_np(4_000_000, 80_000, 0x6772616D + 4)
    haps = synth.indel_haplotypes(ref, sites, 4, 11)
    bases, offs = synth.sample_reads(haps, 20000, 150, 12)
    got, ref_r = _check(prg, 8, bases, offs, what="config4-shape", threads=os.cpu_count())
    assert ref_r.stats[4] >= 20000


def test_packed_host_batch(built_lib):
    """gq_map_batch_packed (2-bit packed host buffers, no pack kernel) == gq_map_batch on the same reads == oracle,
    including skipped (empty) reads, reads shorter than k and several pipelined slices."""
    from gramtools_b200 import pack_reads
    prg = synth.make_nested_prg(5, 300, 21)
    bases, offs = _reads_for(prg, 20000, 45, 21, garbage=0.05, n_frac=0.03)
    seeds = master_seeds(42, offs.size - 1)
    o = Oracle(prg, 5)
    o.map(bases, offs, seeds, threads=os.cpu_count())
    ref = o.result()
    packed, word_off, ln = pack_reads(bases, offs)
    for options in ({}, {"chunk_reads": 4096, "tail_chunk_reads": 1024}):
        idx = QuasimapIndex(prg, 5)
        for k, v in options.items():
            idx.set_option(k, v)
        idx.map_batch_packed(packed, word_off, ln, seeds)
        assert_parity(gpu_result(idx), ref, f"packed{options}")
        assert idx.run_info()["h2d_bytes"] < 0.5 * (bases.size + 12 * (offs.size - 1))
        idx.close()


def test_single_rank_communicator(built_lib):
    """The library's own NCCL path with one rank: communicator from gq_comm_unique_id / gq_comm_init, in-place
    all-reduce of counters, stats and sparse groups — a no-op on the values."""
    from gramtools_b200 import comm_unique_id
    prg = synth.make_nested_prg(3, 250, 31)
    bases, offs = _reads_for(prg, 3000, 30, 31)
    seeds = master_seeds(42, offs.size - 1)
    idx = QuasimapIndex(prg, 4)
    idx.comm_init(comm_unique_id(), 0, 1)
    idx.map_batch(bases, offs, seeds)
    before = gpu_result(idx)
    idx.coverage_allreduce()
    after = gpu_result(idx)
    assert_parity(after, before, "1-rank allreduce")
    o = Oracle(prg, 4)
    o.map(bases, offs, seeds)
    assert_parity(after, o.result(), "1-rank allreduce vs oracle")


def test_two_gpus_one_process_allreduce(built_lib):
    """Reads sharded over two GPUs of this process (index built once, cloned), one NCCL exchange at the end:
    every handle then holds the job's totals — dense counters, stats and merged multi-allele groups — equal to the
    oracle's single-process result. Needs two devices (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from gramtools_b200.distributed import shard_bounds
    prg = synth.make_nested_prg(6, 300, 41)
    bases, offs = _reads_for(prg, 20000, 40, 41)
    n = offs.size - 1
    seeds = master_seeds(42, n)
    a = QuasimapIndex(prg, 5, device=0)
    handles = [a, a.clone(1)]
    QuasimapIndex.comm_init_all(handles)
    for r, h in enumerate(handles):
        lo, hi = shard_bounds(n, r, 2)
        h.map_batch(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], seeds[lo:hi])
    QuasimapIndex.coverage_allreduce_all(handles)
    o = Oracle(prg, 5)
    o.map(bases, offs, seeds, threads=os.cpu_count())
    ref = o.result()
    for h in handles:
        al, pb, st = h.coverage()
        assert np.array_equal(al, ref.allele_sum) and np.array_equal(pb, ref.per_base)
        assert np.array_equal(h.grouped(), ref.grouped)
        assert [st.all_reads_count, st.skipped_reads_count, st.missing_kmer_reads_count, st.no_extension_reads_count,
                st.exact_mapped_reads_count] == ref.stats


def test_gpu_suffix_array(built_lib):
    """The GPU suffix array (prefix doubling, sa_gpu.cu — what gq_index_build uses) against the host SA-IS of the
    test emulation: random SNP / indel / nested PRGs, long exact repeats (many doubling rounds), a periodic text,
    a one-symbol PRG."""
    from common import Emu
    from gramtools_b200 import suffix_array
    rng = np.random.default_rng(3)
    rep = rng.integers(1, 5, size=3000).astype(np.uint32)
    prgs = {
        "snp": synth.make_snp_prg(200_000, 5_000, 1)[0],
        "indel": synth.make_indel_prg(50_000, 2_000, 2),
        "nested": synth.make_nested_prg(20, 800, 3),
        "nested, coinciding alleles": synth.make_nested_prg(6, 500, 4, distinct=False),
        "repeats": np.concatenate([rep, rng.integers(1, 5, size=100).astype(np.uint32), rep, rep[:1500], rep]),
        "periodic": np.tile(np.asarray([1, 2, 3], dtype=np.uint32), 4000),
        "one symbol": np.asarray([3], dtype=np.uint32),
        "run": np.full(5000, 2, dtype=np.uint32),
    }
    for name, prg in prgs.items():
        prg = np.ascontiguousarray(prg, dtype=np.uint32)
        sa, rounds = suffix_array(prg)
        want = Emu(prg, 1 if prg.size < 4 else 3).sa()
        assert np.array_equal(sa, want), name
        assert rounds >= 1
