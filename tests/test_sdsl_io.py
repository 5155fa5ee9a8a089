"""The k-mer index as gram_dir files — kmers, kmers_stats, sa_intervals, paths (dump.cpp:27-141, load.cpp:11-173) —
in sdsl::int_vector serialisation. PARITY UNPINNED against sdsl itself (not available here, no fixture in the
reference): the golden bytes below are written out by hand from sdsl-lite 2.1.1's int_vector::serialize (bit length
as uint64 LE, a width byte for int_vector<0> only, elements packed LSB first into zero-padded 64-bit words)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from common import Emu, Oracle, assert_parity, emu_lib
from gramtools_b200 import encode_reads, master_seeds, synth


def _write(path, values, width, fixed):
    lib = emu_lib()
    v = np.asarray(values, dtype=np.uint64)
    rc = lib.emu_write_int_vector(os.fsencode(path), v.ctypes.data_as(C.POINTER(C.c_uint64)), v.size, width, int(fixed))
    assert rc == 0, lib.emu_last_error().decode()


def _read(path, fixed_width):
    lib = emu_lib()
    out = np.zeros(1 << 16, dtype=np.uint64)
    w = C.c_uint32(0)
    n = lib.emu_read_int_vector(os.fsencode(path), fixed_width, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size, C.byref(w))
    assert n >= 0, lib.emu_last_error().decode()
    return out[:n].tolist(), w.value


def test_int_vector_golden_bytes(tmp_path):
    # int_vector<3> [1,2,3,4]: 12 bits; 1 | 2<<3 | 3<<6 | 4<<9 = 0x8D1
    p = str(tmp_path / "v3")
    _write(p, [1, 2, 3, 4], 3, True)
    assert open(p, "rb").read() == struct.pack("<QQ", 12, 0x8D1)
    assert _read(p, 3) == ([1, 2, 3, 4], 3)
    # int_vector<> of width 7: [100, 27, 127]: 21 bits, width byte 7; 100 | 27<<7 | 127<<14
    p = str(tmp_path / "v0")
    _write(p, [100, 27, 127], 7, False)
    assert open(p, "rb").read() == struct.pack("<QBQ", 21, 7, 100 | (27 << 7) | (127 << 14))
    assert _read(p, 0) == ([100, 27, 127], 7)
    # elements straddling a word boundary: 5 x 13 bits = 65 bits -> two words
    vals = [8191, 1, 4097, 77, 8000]
    p = str(tmp_path / "v13")
    _write(p, vals, 13, False)
    big = sum(v << (13 * i) for i, v in enumerate(vals))
    assert open(p, "rb").read() == struct.pack("<QBQQ", 65, 13, big & (2 ** 64 - 1), big >> 64)
    assert _read(p, 0) == (vals, 13)
    # empty vector: header only
    p = str(tmp_path / "empty")
    _write(p, [], 1, False)
    assert open(p, "rb").read() == struct.pack("<QB", 0, 1)
    assert _read(p, 0) == ([], 1)


@pytest.mark.parametrize("kind", ["snp", "nested", "indel"])
def test_kmer_index_round_trip(tmp_path, kind):
    """dump -> load gives the same index (digest over every array; the loaded index lays the paths out in k-mer order,
    the builder in the order of its subtree tasks, hence the layout-free digest), and the file contents are what
    dump.cpp describes: k codes per k-mer, per-k-mer state counts + path lengths, interval pairs, (site, allele + 1)."""
    prg, k = {"snp": (synth.make_snp_prg(3000, 120, 5)[0], 5), "nested": (synth.make_nested_prg(6, 400, 2), 4),
              "indel": (synth.make_indel_prg(2500, 120, 7), 6)}[kind]
    e = Emu(prg, k)
    e.kmer_index_dump(str(tmp_path))
    kmers, w = _read(str(tmp_path / "kmers"), 3)
    stats, _ = _read(str(tmp_path / "kmers_stats"), 0)
    sa_iv, _ = _read(str(tmp_path / "sa_intervals"), 0)
    paths, _ = _read(str(tmp_path / "paths"), 0)
    assert len(kmers) % k == 0 and set(kmers) <= {1, 2, 3, 4}
    n_kmers = len(kmers) // k
    n_states, i, n_path = 0, 0, 0
    for _ in range(n_kmers):
        ns = stats[i]
        assert ns >= 1
        n_path += sum(stats[i + 1:i + 1 + ns])
        n_states += ns
        i += 1 + ns
    assert i == len(stats) and n_states == e.n_kmer_states and len(sa_iv) == 2 * n_states and len(paths) == 2 * n_path
    assert all(sa_iv[2 * j] <= sa_iv[2 * j + 1] < e.sa_size for j in range(n_states))
    assert all(m >= 5 for m in paths[0::2])
    e2 = Emu(prg, k, kmer_index_dir=str(tmp_path))
    assert e2.index_digest(layout_free=True) == e.index_digest(layout_free=True)
    d3 = tmp_path / "again"  # and a loaded index dumps the same four files
    d3.mkdir()
    e2.kmer_index_dump(str(d3))
    for name in ("kmers", "kmers_stats", "sa_intervals", "paths"):
        assert open(d3 / name, "rb").read() == open(tmp_path / name, "rb").read(), name
    # k-mers in another order in the file (the reference writes its hash map's order): same index
    order = np.random.default_rng(1).permutation(n_kmers)
    starts, i = [], 0
    for _ in range(n_kmers):
        starts.append(i)
        i += 1 + stats[i]
    st_at = np.cumsum([0] + [stats[s] for s in starts])
    pa_at = np.cumsum([0] + [2 * sum(stats[s + 1:s + 1 + stats[s]]) for s in starts])
    d2 = tmp_path / "shuffled"
    d2.mkdir()
    _write(str(d2 / "kmers"), [x for q in order for x in kmers[q * k:(q + 1) * k]], 3, True)
    _write(str(d2 / "kmers_stats"), [x for q in order for x in stats[starts[q]:starts[q] + 1 + stats[starts[q]]]], 8, False)
    _write(str(d2 / "sa_intervals"), [x for q in order for x in sa_iv[2 * st_at[q]:2 * st_at[q + 1]]], 32, False)
    _write(str(d2 / "paths"), [x for q in order for x in paths[pa_at[q]:pa_at[q + 1]]], 32, False)
    assert Emu(prg, k, kmer_index_dir=str(d2)).index_digest(layout_free=True) == e.index_digest(layout_free=True)


def test_index_from_files_maps_like_the_oracle(tmp_path):
    prg, k = synth.make_nested_prg(5, 300, 9), 4
    Emu(prg, k).kmer_index_dump(str(tmp_path))
    e = Emu(prg, k, kmer_index_dir=str(tmp_path))
    rng = np.random.default_rng(2)
    haps = [synth.random_haplotype(prg, rng) for _ in range(3)]
    bases, offs = synth.sample_reads(haps, 200, 40, 3)
    seeds = master_seeds(42, offs.size - 1)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    e.map(bases, offs, seeds)
    assert_parity(e.result(), o.result(), "index loaded from gram_dir files")


def test_malformed_files_rejected(tmp_path):
    prg, k = synth.make_snp_prg(800, 30, 5)[0], 4
    Emu(prg, k).kmer_index_dump(str(tmp_path))
    raw = open(tmp_path / "sa_intervals", "rb").read()
    open(tmp_path / "sa_intervals", "wb").write(raw[:-8])  # truncated
    with pytest.raises(RuntimeError):
        Emu(prg, k, kmer_index_dir=str(tmp_path))
    Emu(prg, k).kmer_index_dump(str(tmp_path))
    with pytest.raises(RuntimeError):
        Emu(prg, k + 1, kmer_index_dir=str(tmp_path))  # files of another kmer_size
    with pytest.raises(RuntimeError):
        Emu(prg, k, kmer_index_dir=str(tmp_path / "missing"))


def test_whole_index_file_round_trip(tmp_path):
    """gq_index (this back-end's own format: every array of the flat index + a checksum): save -> load gives the same
    bytes in memory (full digest), maps like the oracle, and damaged files are refused."""
    prg, k = synth.make_nested_prg(6, 350, 4), 5
    e = Emu(prg, k)
    path = str(tmp_path / "gq_index")
    e.index_save(path)
    e2 = Emu(None, 0, index_file=path)
    assert e2.index_digest() == e.index_digest()
    assert (e2.n_sites, e2.n_alleles, e2.n_per_base, e2.sa_size, e2.n_kmer_states) == \
           (e.n_sites, e.n_alleles, e.n_per_base, e.sa_size, e.n_kmer_states)
    rng = np.random.default_rng(5)
    haps = [synth.random_haplotype(prg, rng) for _ in range(3)]
    bases, offs = synth.sample_reads(haps, 200, 50, 6)
    seeds = master_seeds(7, offs.size - 1)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    e2.map(bases, offs, seeds)
    assert_parity(e2.result(), o.result(), "index loaded from a gq_index file")
    raw = open(path, "rb").read()
    for name, data in (("truncated", raw[:len(raw) // 2]), ("flipped", raw[:1000] + bytes([raw[1000] ^ 1]) + raw[1001:]),
                       ("magic", b"NOTANIDX" + raw[8:]), ("trailing", raw + b"\0" * 8), ("empty", b"")):
        bad = str(tmp_path / name)
        open(bad, "wb").write(data)
        with pytest.raises(RuntimeError):
            Emu(None, 0, index_file=bad)
    with pytest.raises(RuntimeError):
        Emu(None, 0, index_file=str(tmp_path / "missing"))


def test_int_vector_random_round_trips(tmp_path):
    """Every width 1..64, random lengths and values: the bytes are the big-integer packing (element i at bit i * width,
    LSB first) and reading gives the values back."""
    rng = np.random.default_rng(9)
    for width in list(range(1, 65)):
        n = int(rng.integers(0, 40))
        vals = [int(rng.integers(0, 2 ** min(width, 63))) | ((int(rng.integers(0, 2)) << 63) if width == 64 else 0)
                for _ in range(n)]
        for fixed in (False, True):
            p = str(tmp_path / f"w{width}{'f' if fixed else 'v'}")
            _write(p, vals, width, fixed)
            raw = open(p, "rb").read()
            big = sum(v << (width * i) for i, v in enumerate(vals))
            n_words = (n * width + 63) // 64
            want = struct.pack("<Q", n * width) + (b"" if fixed else bytes([width])) + big.to_bytes(8 * n_words, "little")
            assert raw == want, (width, fixed)
            assert _read(p, width if fixed else 0) == (vals, width)


# ---- the k-mer indexes of the reference's own dump / load tests (tests/build/kmer_index/test_dump_and_load.cpp) ----
# Each literal is (k-mer bases 1..4, [(lo, hi, traversed [(site, allele)...], traversing [site...])...]); the expected file
# contents follow dump.cpp:27-141 (k base codes per k-mer; per k-mer the state count then one path length per state;
# (lo, hi) per state; (site, allele + 1) per traversed locus then (site, 0) per traversing one).
_UNK = 0xFFFFFFFF
REFERENCE_DUMP_CASES = {
    # DumpKmers.GivenTwoKmers_CorrectAllKmersStructure (:14-28): all_kmers == {1,2,3,4, 2,4,3,4} in either order
    "two_kmers": [((1, 2, 3, 4), [(1, 1, [], [])]), ((2, 4, 3, 4), [(2, 2, [], [])])],
    # DumpAndLoadIndex.SearchStatesWithNoVariants (:82-98)
    "no_variants": [((4, 4, 4, 4), [(20000, 22000, [], []), (52, 53, [], []), (62, 63, [], [])])],
    # DumpAndLoadIndex.SearchStateVariantsWithLargeIndices (:100-126): > 1 billion sites / alleles
    "large_indices": [((1, 2, 3, 4), [(6, 6, [(1200000000, 0)], []), (7, 42, [(5, 1200000000)], [])])],
    # DumpAndLoadIndex.TwoPathsWithMultipleElements (:128-144)
    "two_paths": [((1, 2, 3, 4), [(6, 6, [(5, 1)], []), (7, 42, [(7, 3), (5, 2)], [9])])],
    # DumpAndLoadIndex.TwoKmersWithMultipleSearchStates (:146-182)
    "two_kmers_states": [((1, 2, 3, 4), [(6, 6, [(5, 0)], []), (7, 7, [(5, 1)], []), (8, 8, [(5, 1)], [])]),
                         ((2, 4, 3, 4), [(9, 10, [], []), (11, 11, [(5, 1), (7, 1)], [])])],
    # DumpAndLoadIndex.WithTraversingPaths (:184-208)
    "traversing": [((1, 2, 3, 4), [(6, 6, [(5, 0)], [7]), (7, 7, [(5, 1)], []), (8, 8, [(5, 1)], [11, 9])])],
}


def _code(bases):
    return sum((b - 1) << (2 * j) for j, b in enumerate(bases))


@pytest.mark.parametrize("name", sorted(REFERENCE_DUMP_CASES))
def test_reference_dump_and_load_literals(tmp_path, name):
    case = REFERENCE_DUMP_CASES[name]
    k = len(case[0][0])
    prg = synth.make_snp_prg(30000, 10, 1)[0]  # only its size matters: the literals' SA indices go up to 22000
    e = Emu(prg, k)
    words = []
    for bases, states in case:
        for lo, hi, trav, ing in states:
            words += [_code(bases), lo, hi, len(trav), len(ing)] + [x for loc in trav for x in loc] + [x for s in ing for x in (s, _UNK)]
    e.set_kmer_index(words)
    e.kmer_index_dump(str(tmp_path))
    kmers, _ = _read(str(tmp_path / "kmers"), 3)
    stats, _ = _read(str(tmp_path / "kmers_stats"), 0)
    sa_iv, _ = _read(str(tmp_path / "sa_intervals"), 0)
    paths, _ = _read(str(tmp_path / "paths"), 0)
    by_code = sorted(case, key=lambda c: _code(c[0]))  # this writer's order (the reference's is its hash map's)
    assert kmers == [b for bases, _ in by_code for b in bases]
    assert stats == [x for _, states in by_code for x in [len(states)] + [len(t) + len(g) for _, _, t, g in states]]
    assert sa_iv == [x for _, states in by_code for lo, hi, _, _ in states for x in (lo, hi)]
    assert paths == [x for _, states in by_code for _, _, t, g in states
                     for x in [y for site, al in t for y in (site, al + 1)] + [y for site in g for y in (site, 0)]]
    # load gives the index back (DumpAndLoadIndex: EXPECT_EQ(load(dump(index)), index))
    loaded = Emu(prg, k, kmer_index_dir=str(tmp_path))
    want = []
    for bases, states in by_code:
        for lo, hi, trav, ing in states:
            want += [_code(bases), lo, hi, len(trav), len(ing)] + [x for loc in trav for x in loc] + [x for s in ing for x in (s, _UNK)]
    assert loaded.kmer_states() == want


def test_reference_deserialize_next_stats_layout(tmp_path):
    """DeserializeNextStats (:34-72): kmers_stats {3, 1, 42, 7, 2, 11, 33} = a k-mer with three states of path lengths
    1, 42, 7, then one with two states of path lengths 11, 33 — the layout this reader consumes."""
    p = str(tmp_path / "kmers_stats")
    _write(p, [3, 1, 42, 7, 2, 11, 33], 6, False)
    vals, width = _read(p, 0)
    assert (vals, width) == ([3, 1, 42, 7, 2, 11, 33], 6)
    i, got = 0, []
    while i < len(vals):
        got.append((vals[i], vals[i + 1:i + 1 + vals[i]]))
        i += 1 + vals[i]
    assert got == [(3, [1, 42, 7]), (2, [11, 33])]


def test_dumped_vectors_are_those_of_the_oracles_index(tmp_path):
    """gram build's four k-mer index files for a real PRG: the vectors hold what dump.cpp:27-141 prescribes for the
    ORACLE's k-mer index — every k-mer's states in the reference's list order (k-mers by ascending code here, in hash-map
    order in the reference: load.cpp takes either)."""
    for n, (prg, k) in enumerate(((synth.make_snp_prg(3000, 120, 6)[0], 5), (synth.make_nested_prg(6, 250, 21), 4))):
        d = tmp_path / f"gram{n}"
        d.mkdir()
        Emu(prg, k).kmer_index_dump(str(d))
        words = Oracle(prg, k).kmer_states()
        by_kmer, i = {}, 0   # code -> [(lo, hi, [(site, allele)...], [site...])]
        while i < len(words):
            code, lo, hi, nt, ng = words[i:i + 5]
            trav = [(words[i + 5 + 2 * j], words[i + 6 + 2 * j]) for j in range(nt)]
            ing = [words[i + 5 + 2 * nt + 2 * j] for j in range(ng)]
            by_kmer.setdefault(code, []).append((lo, hi, trav, ing))
            i += 5 + 2 * nt + 2 * ng
        codes = sorted(by_kmer)
        kmers, _ = _read(str(d / "kmers"), 3)
        stats, _ = _read(str(d / "kmers_stats"), 0)
        sa_iv, _ = _read(str(d / "sa_intervals"), 0)
        paths, _ = _read(str(d / "paths"), 0)
        assert kmers == [((c >> (2 * j)) & 3) + 1 for c in codes for j in range(k)]
        assert stats == [x for c in codes for x in [len(by_kmer[c])] + [len(t) + len(g) for _, _, t, g in by_kmer[c]]]
        assert sa_iv == [x for c in codes for lo, hi, _, _ in by_kmer[c] for x in (lo, hi)]
        assert paths == [x for c in codes for _, _, t, g in by_kmer[c]
                         for x in [y for site, al in t for y in (site, al + 1)] + [y for site in g for y in (site, 0)]]
