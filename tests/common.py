"""Shared test plumbing: ctypes bindings of the CPU oracle (oracle/) and of the test-only host
emulation of the device functions (tests/emu), canonicalisation of per-strand state records, and the
parity comparator used by both the CPU and the GPU suites."""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _make(target_dir):
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", target_dir], check=True, env=env, stdout=subprocess.DEVNULL)


_oracle = None
_emu = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        p = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(p):
            _make(os.path.join(ROOT, "oracle"))
        lib = C.CDLL(p)
        lib.gqo_new.restype = C.c_void_p
        lib.gqo_new.argtypes = [u32p, C.c_uint64, C.c_uint32]
        lib.gqo_free.argtypes = [C.c_void_p]
        lib.gqo_last_error.restype = C.c_char_p
        lib.gqo_master_seeds.argtypes = [C.c_uint32, C.c_uint64, u32p]
        lib.gqo_sizes.argtypes = [C.c_void_p, u64p]
        lib.gqo_kmer_states.argtypes = [C.c_void_p, u32p]
        lib.gqo_kmer_states.restype = C.c_uint64
        lib.gqo_map.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u32p, C.c_int, C.c_int, C.c_int]
        lib.gqo_map_forward.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u32p]
        lib.gqo_last_seconds.argtypes = [C.c_void_p]
        lib.gqo_last_seconds.restype = C.c_double
        lib.gqo_status.argtypes = [C.c_void_p, u8p]
        lib.gqo_states_size.argtypes = [C.c_void_p]
        lib.gqo_states_size.restype = C.c_uint64
        lib.gqo_states.argtypes = [C.c_void_p, u64p, u32p, u32p]
        lib.gqo_allele_sum.argtypes = [C.c_void_p, u16p]
        lib.gqo_per_base.argtypes = [C.c_void_p, u16p]
        lib.gqo_grouped.argtypes = [C.c_void_p, u32p]
        lib.gqo_grouped.restype = C.c_uint64
        lib.gqo_stats.argtypes = [C.c_void_p, u64p]
        lib.gqo_events.argtypes = [C.c_void_p, u64p]
        _oracle = lib
    return _oracle


def emu_lib():
    global _emu
    if _emu is None:
        p = os.path.join(ROOT, "tests", "_build", "libgq_emu.so")
        _make(os.path.join(ROOT, "tests", "emu"))
        lib = C.CDLL(p)
        lib.emu_new.restype = C.c_void_p
        lib.emu_new.argtypes = [u32p, C.c_uint64, C.c_uint32]
        lib.emu_new_from.restype = C.c_void_p
        lib.emu_new_from.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_char_p]
        lib.emu_kmer_index_dump.argtypes = [C.c_void_p, C.c_char_p]
        lib.emu_index_save.argtypes = [C.c_void_p, C.c_char_p]
        lib.emu_set_kmer_index.argtypes = [C.c_void_p, u32p, C.c_uint64]
        lib.emu_read_file.argtypes = [C.c_char_p, C.c_int, C.c_uint64, u64p]
        lib.emu_load.restype = C.c_void_p
        lib.emu_load.argtypes = [C.c_char_p]
        lib.emu_write_int_vector.argtypes = [C.c_char_p, u64p, C.c_uint64, C.c_uint32, C.c_int]
        lib.emu_read_int_vector.argtypes = [C.c_char_p, C.c_uint32, u64p, C.c_uint64, C.POINTER(C.c_uint32)]
        lib.emu_read_int_vector.restype = C.c_int64
        lib.emu_free.argtypes = [C.c_void_p]
        lib.emu_last_error.restype = C.c_char_p
        lib.emu_sizes.argtypes = [C.c_void_p, u64p]
        lib.emu_kmer_states.argtypes = [C.c_void_p, u32p]
        lib.emu_kmer_states.restype = C.c_uint64
        lib.emu_sa.argtypes = [C.c_void_p, u32p]
        lib.emu_sais64_agrees.argtypes = [C.POINTER(C.c_int32), C.c_uint64, C.c_int32]
        lib.emu_map.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64, u32p, C.c_uint32]
        lib.emu_reruns.argtypes = [C.c_void_p]
        lib.emu_reruns.restype = C.c_uint64
        lib.emu_status.argtypes = [C.c_void_p, u8p]
        lib.emu_states_size.argtypes = [C.c_void_p]
        lib.emu_states_size.restype = C.c_uint64
        lib.emu_states.argtypes = [C.c_void_p, u64p, u32p, u32p]
        lib.emu_allele_sum.argtypes = [C.c_void_p, u16p]
        lib.emu_per_base.argtypes = [C.c_void_p, u16p]
        lib.emu_grouped.argtypes = [C.c_void_p, u32p]
        lib.emu_grouped.restype = C.c_uint64
        lib.emu_stats.argtypes = [C.c_void_p, u64p]
        lib.emu_index_check.argtypes = [C.c_void_p]
        lib.emu_index_digest.argtypes = [C.c_void_p]
        lib.emu_index_digest.restype = C.c_uint64
        lib.emu_index_digest2.argtypes = [C.c_void_p, C.c_int]
        lib.emu_index_digest2.restype = C.c_uint64
        lib.emu_path_counters.argtypes = [u64p, C.c_int]
        lib.emu_force_general.argtypes = [C.c_int]
        lib.emu_set_gtab_cap.argtypes = [C.c_void_p, C.c_uint32]
        lib.emu_track.argtypes = [C.c_int]
        lib.emu_track_ranges.restype = C.c_int
        lib.emu_track_range_name.argtypes = [C.c_int]
        lib.emu_track_range_name.restype = C.c_char_p
        lib.emu_track_read.argtypes = [u64p, u64p]
        lib.emu_counters_raw.argtypes = [C.c_void_p, u32p]
        lib.emu_groups_raw.argtypes = [C.c_void_p, u32p]
        lib.emu_groups_raw.restype = C.c_uint64
        _emu = lib
    return _emu


@dataclass
class Result:
    status: np.ndarray
    state_off: np.ndarray
    state_count: np.ndarray
    state_words: np.ndarray
    allele_sum: np.ndarray
    per_base: np.ndarray
    grouped: np.ndarray
    stats: list
    extra: dict = field(default_factory=dict)


def canonical_strand(words, count):
    """Sort the `count` records [lo,hi,nt,ng,pairs...] of one strand lexicographically."""
    recs, i = [], 0
    words = [int(x) for x in words]
    for _ in range(int(count)):
        n = 4 + 2 * words[i + 2] + 2 * words[i + 3]
        recs.append(tuple(words[i:i + n]))
        i += n
    assert i == len(words), "record stream does not match its count"
    recs.sort()
    return recs


class Oracle:
    def __init__(self, prg, k):
        self.lib = oracle_lib()
        prg = np.ascontiguousarray(prg, dtype=np.uint32)
        self.h = self.lib.gqo_new(_ptr(prg, C.c_uint32), prg.size, k)
        if not self.h:
            raise RuntimeError(self.lib.gqo_last_error().decode())
        s = np.zeros(6, dtype=np.uint64)
        self.lib.gqo_sizes(self.h, _ptr(s, C.c_uint64))
        self.n_sites, self.n_alleles, self.n_per_base, self.is_nested, self.sa_size, self.n_kmer_states = [int(x) for x in s]

    def kmer_states(self):
        """The oracle's k-mer index: [code, lo, hi, nt, ng, (site, allele) * nt, (site, ~0) * ng] records, k-mers by
        ascending code, states in the reference's order."""
        n = self.lib.gqo_kmer_states(self.h, None)
        w = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.gqo_kmer_states(self.h, _ptr(w, C.c_uint32))
        return w[:n].tolist()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.gqo_free(self.h)
            self.h = None

    def map(self, bases, offsets, seeds, threads=1, want_states=True, count_events=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = offsets.size - 1
        rc = self.lib.gqo_map(self.h, _ptr(bases, C.c_uint8), _ptr(offsets, C.c_uint64), n, _ptr(seeds, C.c_uint32),
                              threads, int(want_states), int(count_events))
        if rc != 0:
            raise RuntimeError(self.lib.gqo_last_error().decode())
        self.n_reads = n
        return self.lib.gqo_last_seconds(self.h)

    def map_forward(self, bases, offsets, seeds):
        """quasimap_read on the forward strand only (the reference's test helper, test_resources.cpp:48-56)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = offsets.size - 1
        if self.lib.gqo_map_forward(self.h, _ptr(bases, C.c_uint8), _ptr(offsets, C.c_uint64), n, _ptr(seeds, C.c_uint32)) != 0:
            raise RuntimeError(self.lib.gqo_last_error().decode())
        self.n_reads = n

    def result(self, want_states=True):
        n = self.n_reads
        status = np.zeros(2 * n, dtype=np.uint8)
        self.lib.gqo_status(self.h, _ptr(status, C.c_uint8))
        off = np.zeros(2 * n + 1, dtype=np.uint64)
        cnt = np.zeros(2 * n, dtype=np.uint32)
        words = np.zeros(1, dtype=np.uint32)
        if want_states:
            nw = self.lib.gqo_states_size(self.h)
            words = np.zeros(max(nw, 1), dtype=np.uint32)
            self.lib.gqo_states(self.h, _ptr(off, C.c_uint64), _ptr(cnt, C.c_uint32), _ptr(words, C.c_uint32))
            words = words[:nw]
        a = np.zeros(max(self.n_alleles, 1), dtype=np.uint16)
        self.lib.gqo_allele_sum(self.h, _ptr(a, C.c_uint16))
        p = np.zeros(max(self.n_per_base, 1), dtype=np.uint16)
        self.lib.gqo_per_base(self.h, _ptr(p, C.c_uint16))
        ng = self.lib.gqo_grouped(self.h, None)
        g = np.zeros(max(ng, 1), dtype=np.uint32)
        self.lib.gqo_grouped(self.h, _ptr(g, C.c_uint32))
        st = np.zeros(5, dtype=np.uint64)
        self.lib.gqo_stats(self.h, _ptr(st, C.c_uint64))
        return Result(status, off, cnt, words, a[:self.n_alleles], p[:self.n_per_base], g[:ng], [int(x) for x in st])

    def events(self):
        e = np.zeros(7, dtype=np.uint64)
        self.lib.gqo_events(self.h, _ptr(e, C.c_uint64))
        return dict(zip(["q_rank", "w_marker", "q_sa", "q_node", "a_cov", "strands", "bases"], [int(x) for x in e]))


class Emu:
    def __init__(self, prg, k, kmer_index_dir=None, index_file=None):
        self.lib = emu_lib()
        if index_file is not None:  # a whole index written by index_save()
            self.h = self.lib.emu_load(os.fsencode(index_file))
        else:
            prg = np.ascontiguousarray(prg, dtype=np.uint32)
            self.h = self.lib.emu_new_from(_ptr(prg, C.c_uint32), prg.size, k,
                                           os.fsencode(kmer_index_dir) if kmer_index_dir else None)
        if not self.h:
            raise RuntimeError(self.lib.emu_last_error().decode())
        s = np.zeros(6, dtype=np.uint64)
        self.lib.emu_sizes(self.h, _ptr(s, C.c_uint64))
        self.n_sites, self.n_alleles, self.n_per_base, self.is_nested, self.sa_size, self.n_kmer_states = [int(x) for x in s]

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.emu_free(self.h)
            self.h = None

    def sa(self):
        out = np.zeros(self.sa_size, dtype=np.uint32)
        self.lib.emu_sa(self.h, _ptr(out, C.c_uint32))
        return out

    def map(self, bases, offsets, seeds, arena_words=256):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = offsets.size - 1
        rc = self.lib.emu_map(self.h, _ptr(bases, C.c_uint8), _ptr(offsets, C.c_uint64), n, _ptr(seeds, C.c_uint32),
                              arena_words)
        if rc != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())
        self.n_reads = n

    ROUTES = ["seed_finished", "too_wide", "seed_states", "-", "multi_finisher", "unresolved_jump", "wide_entry",
              "fast_finished", "general", "many_seed_states"]

    def routes(self, reset=True):
        """Which route strands took since the last reset (process-wide counters of the emulation)."""
        c = np.zeros(32, dtype=np.uint64)
        self.lib.emu_path_counters(_ptr(c, C.c_uint64), int(reset))
        return {k: int(v) for k, v in zip(self.ROUTES, c) if k != "-"}

    def set_kmer_index(self, words):
        """Replace the k-mer index by records [code, lo, hi, nt, ng, (site, allele) * nt, (site, _) * ng]."""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        if self.lib.emu_set_kmer_index(self.h, _ptr(w, C.c_uint32), w.size) != 0:
            raise RuntimeError("bad k-mer records")

    def kmer_states(self):
        n = self.lib.emu_kmer_states(self.h, None)
        w = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.emu_kmer_states(self.h, _ptr(w, C.c_uint32))
        return w[:n].tolist()

    def index_save(self, path):
        """The whole flat index as one checksummed file (host_index_save)."""
        if self.lib.emu_index_save(self.h, os.fsencode(path)) != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def kmer_index_dump(self, gram_dir):
        """kmers / kmers_stats / sa_intervals / paths in the reference's sdsl format (dump.cpp:27-141)."""
        if self.lib.emu_kmer_index_dump(self.h, os.fsencode(gram_dir)) != 0:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def index_digest(self, layout_free=False):
        """FNV-1a over every array of the flat index (layout_free: k-mer states with their path words inline, so the
        order of the paths in the pool does not matter)."""
        return int(self.lib.emu_index_digest2(self.h, 1 if layout_free else 0))

    def index_check(self):
        if self.lib.emu_index_check(self.h) != 0:
            raise AssertionError(self.lib.emu_last_error().decode())

    def force_general(self, on):
        self.lib.emu_force_general(int(on))

    KERNELS = ["seed_kernel", "verify_kernel", "text_kernel", "search_kernel", "classify_kernel", "coverage_kernel"]

    def track(self, on):
        """Byte model: record the distinct 32 B sectors every unit of work touches (resets the totals)."""
        self.lib.emu_track(int(on))

    def tracked(self):
        """-> {kernel: {structure: bytes}}, {kernel: units of work} since track(True)"""
        n = self.lib.emu_track_ranges()
        b = np.zeros(6 * max(n, 1), dtype=np.uint64)
        u = np.zeros(6, dtype=np.uint64)
        self.lib.emu_track_read(_ptr(b, C.c_uint64), _ptr(u, C.c_uint64))
        names = [self.lib.emu_track_range_name(i).decode() for i in range(n)]
        out = {}
        for k, kn in enumerate(self.KERNELS):
            out[kn] = {names[i]: int(b[k * n + i]) for i in range(n) if b[k * n + i]}
        return out, {kn: int(u[k]) for k, kn in enumerate(self.KERNELS)}

    def set_gtab_cap(self, cap):
        """Shrink the multi-allele group table (power of two) so that its growth path runs."""
        self.lib.emu_set_gtab_cap(self.h, int(cap))

    def counters_raw(self):
        out = np.zeros(max(2 * self.n_alleles + self.n_per_base, 1), dtype=np.uint32)
        self.lib.emu_counters_raw(self.h, _ptr(out, C.c_uint32))
        return out[:2 * self.n_alleles + self.n_per_base]

    def groups_raw(self):
        n = self.lib.emu_groups_raw(self.h, None)
        w = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.emu_groups_raw(self.h, _ptr(w, C.c_uint32))
        return w[:n]

    def stats(self):
        st = np.zeros(5, dtype=np.uint64)
        self.lib.emu_stats(self.h, _ptr(st, C.c_uint64))
        return st

    def result(self):
        n = self.n_reads
        status = np.zeros(2 * n, dtype=np.uint8)
        self.lib.emu_status(self.h, _ptr(status, C.c_uint8))
        nw = self.lib.emu_states_size(self.h)
        off = np.zeros(2 * n + 1, dtype=np.uint64)
        cnt = np.zeros(2 * n, dtype=np.uint32)
        words = np.zeros(max(nw, 1), dtype=np.uint32)
        self.lib.emu_states(self.h, _ptr(off, C.c_uint64), _ptr(cnt, C.c_uint32), _ptr(words, C.c_uint32))
        a = np.zeros(max(self.n_alleles, 1), dtype=np.uint16)
        self.lib.emu_allele_sum(self.h, _ptr(a, C.c_uint16))
        p = np.zeros(max(self.n_per_base, 1), dtype=np.uint16)
        self.lib.emu_per_base(self.h, _ptr(p, C.c_uint16))
        ng = self.lib.emu_grouped(self.h, None)
        g = np.zeros(max(ng, 1), dtype=np.uint32)
        self.lib.emu_grouped(self.h, _ptr(g, C.c_uint32))
        st = np.zeros(5, dtype=np.uint64)
        self.lib.emu_stats(self.h, _ptr(st, C.c_uint64))
        return Result(status, off, cnt, words[:nw], a[:self.n_alleles], p[:self.n_per_base], g[:ng], [int(x) for x in st],
                      dict(reruns=int(self.lib.emu_reruns(self.h))))


def gpu_result(idx):
    """Result of a gramtools_b200.QuasimapIndex after map_batch()."""
    status = idx.batch_status()
    off, cnt, words = idx.batch_states()
    a, p, st = idx.coverage()
    g = idx.grouped()
    stats = [st.all_reads_count, st.skipped_reads_count, st.missing_kmer_reads_count, st.no_extension_reads_count,
             st.exact_mapped_reads_count]
    return Result(status, off, cnt, words, a, p, g, stats, idx.run_info())


def assert_parity(got: Result, ref: Result, what="", check_states=True):
    """Bit-exact comparison (integer path: no tolerance)."""
    assert got.stats == ref.stats, f"{what}: stats {got.stats} != {ref.stats}"
    bad = np.nonzero(got.status != ref.status)[0]
    assert bad.size == 0, f"{what}: status differs at strands {bad[:10]}: {got.status[bad[:10]]} vs {ref.status[bad[:10]]}"
    if check_states:
        assert np.array_equal(got.state_count, ref.state_count), \
            f"{what}: state counts differ at strands {np.nonzero(got.state_count != ref.state_count)[0][:10]}"
        same = (got.state_off.size == ref.state_off.size and np.array_equal(got.state_off, ref.state_off)
                and np.array_equal(got.state_words, ref.state_words))
        if not same:  # unordered within a strand: canonicalise strand by strand
            for s in range(got.status.size):
                a = canonical_strand(got.state_words[int(got.state_off[s]):int(got.state_off[s + 1])], got.state_count[s])
                b = canonical_strand(ref.state_words[int(ref.state_off[s]):int(ref.state_off[s + 1])], ref.state_count[s])
                assert a == b, f"{what}: final SearchStates differ for strand {s}:\n got {a}\n ref {b}"
    assert np.array_equal(got.allele_sum, ref.allele_sum), \
        f"{what}: allele_sum differs at {np.nonzero(got.allele_sum != ref.allele_sum)[0][:10]}"
    assert np.array_equal(got.per_base, ref.per_base), \
        f"{what}: per-base coverage differs at {np.nonzero(got.per_base != ref.per_base)[0][:10]}"
    assert np.array_equal(got.grouped, ref.grouped), f"{what}: grouped allele counts differ"


# ---------------------------------------------------------------------------------------------------
# Shared cases (used by the CPU suite through the emulation and by the GPU suite through libgq.so)
# ---------------------------------------------------------------------------------------------------
def numbered_prg(numbered):
    """'gct5c6g6t6ag' -> ints (the notation of the reference's tests, prg_string_to_ints)."""
    prg, num = [], ""
    for ch in numbered:
        if ch.isdigit():
            num += ch
        else:
            if num:
                prg.append(int(num))
                num = ""
            prg.append("acgt".index(ch.lower()) + 1)
    if num:
        prg.append(int(num))
    return np.asarray(prg, dtype=np.uint32)


def bracket_prg(br):
    """'a[c,g[ct,t]a]c' -> ints (nested bracket notation of test_linearised_prg / test_covGraph)."""
    stack, nid, prg = [], 3, []
    for ch in br:
        if ch == "[":
            nid += 2
            stack.append(nid)
            prg.append(nid)
        elif ch == "]":
            prg.append(stack.pop() + 1)
        elif ch == ",":
            prg.append(stack[-1] + 1)
        else:
            prg.append("acgt".index(ch.lower()) + 1)
    return np.asarray(prg, dtype=np.uint32)


REFERENCE_SEEDS = (42, 150, 29, 200)  # test_quasimap.cpp:174-198,240-258,386-404


def reference_test_cases():
    """PRGs + reads of the reference's quasimap tests (test_quasimap.cpp), both strands, with the seeds its
    seed-dependent selection tests use. Yields (name, prg, k, reads, seeds_to_try)."""
    cases = [
        ("gct5c6g6t6ag7t8c8cta", ["agccta", "agtcta", "ctgagtcta", "tagtcta", "tgtcta", "gctc", "tagt", "gagt", "cagc"]),
        ("TAG5Tc6g6T6AG7T8c8cta", ["tagt"] * 8),
        ("gtagtac5gtagtact6t6ta", ["gtagt"] * 8),
        ("ac5gtagtact6t6gggtagt6ta", ["gtagt"] * 4),
        ("tac5gta6gtt6ta", ["tacgt"]),
        ("gcac5t6g6c6ta7t8c8cta", ["accta", "gcact"]),
    ]
    for numbered, reads in cases:
        yield numbered, numbered_prg(numbered), 2, reads, REFERENCE_SEEDS
    nested = ["a[c,g[ct,t]a]c", "t[a[c,g][c,g],]t", "A[[A[CCC,c],t],g]TA", "a[t[tt,t]t,a[at,]a]g[c,g]",
              "[AC,[C,G]]T", "[C,G][C,G]", "A[C,,G]T", "AT[GC[GCC,CCGC],T]TTTT", "AAT[ATAT,AA,]AGG"]
    nreads = ["agtac", "tt", "tacct", "AACCCTA", "CTA", "ATTTTGC", "TT", "AAAGG", "ACT", "CT", "GT", "AT",
              "CGCCTT", "ATTTT", "GCC", "CTTT", "ATAT", "ATAAA", "AATAGG"]
    for br in nested:
        for k in (1, 2):
            yield br, bracket_prg(br), k, nreads, (42,)


def uint16_case():
    """One batch that pushes every kind of counter past 65535: a SNP allele (allele_sum and its single-allele
    group WRAP, allele_sum.cpp:40-41 / grouped_allele_counts.cpp:44-47), a multi-allele group (a read that
    starts inside a site on a suffix shared by two alleles: wraps), and per-base cells (SATURATE at 65535,
    allele_base.cpp:239-241). Returns (prg, k, reads)."""
    rng = np.random.default_rng(12)
    s = lambda a: "".join("?ACGT"[int(x)] for x in a)
    left, mid, right = rng.integers(1, 5, 40), rng.integers(1, 5, 30), rng.integers(1, 5, 40)
    # site 5: SNP A/C; site 7: alleles GGTA / CCTA / T (the first two share the suffix TA)
    prg = np.concatenate([left, [5, 1, 6, 2, 6], mid, [7, 3, 3, 4, 1, 8, 2, 2, 4, 1, 8, 4, 8], right]).astype(np.uint32)
    snp_read = s(left[-12:]) + "C" + s(mid[:12])          # crosses site 5 through allele 1
    shared = "TA" + s(right[:20])                          # starts inside site 7: alleles 0 and 1 both fit
    reads = [snp_read] * 70000 + [shared] * 66000 + [s(mid[-10:]) + "T" + s(right[:15])] * 300
    return prg, 5, reads
