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
