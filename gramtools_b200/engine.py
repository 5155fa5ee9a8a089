"""ctypes binding of libgq.so (include/gq.h).

Host-side mirror of the reference seam ``gram::quasimap_reads`` (quasimap.hpp:29-32): build /
load the index, map batches of encoded reads, fetch ``Coverage`` + ``QuasimapReadsStats``.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class GqError(RuntimeError):
    pass


def lib_path():
    # GQ_LIB: developer override to A/B-test another build of the same library (tools/)
    return os.environ.get("GQ_LIB") or os.path.join(_HERE, "libgq.so")


_lib = None


def load_library():
    """Load libgq.so; fail loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise GqError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
    vp = C.c_void_p
    sig = {
        "gq_index_build": [u32p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(vp)],
        "gq_index_destroy": [vp],
        "gq_index_describe": [vp, C.POINTER(GqLayout)],
        "gq_index_allele_offsets": [vp, u64p],
        "gq_index_per_base_layout": [vp, u64p],
        "gq_map_batch": [vp, u8p, u64p, C.c_uint64, u32p],
        "gq_batch_upload": [vp, u8p, u64p, C.c_uint64, u32p],
        "gq_map_resident": [vp],
        "gq_batch_status": [vp, u8p],
        "gq_batch_states_size": [vp, u64p],
        "gq_batch_states": [vp, u64p, u32p, u32p],
        "gq_coverage_fetch": [vp, u16p, u16p, u64p],
        "gq_coverage_grouped": [vp, u32p, u64p],
        "gq_coverage_reset": [vp],
        "gq_read_depth_stats": [vp, C.POINTER(C.c_double), u64p],
        "gq_coverage_device_ptrs": [vp, C.POINTER(vp), u64p, C.POINTER(vp)],
        "gq_coverage_groups_export": [vp, u32p, u64p],
        "gq_coverage_groups_import": [vp, u32p, C.c_uint64, C.c_int],
        "gq_set_stream": [vp, vp],
        "gq_set_option": [vp, C.c_char_p, C.c_int64],
        "gq_last_run_info": [vp, C.POINTER(C.c_double)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.gq_last_error.restype = C.c_char_p
    _lib = lib
    return lib


class GqLayout(C.Structure):
    _fields_ = [("n_symbols", C.c_uint64), ("sa_size", C.c_uint64), ("kmer_size", C.c_uint32),
                ("n_sites", C.c_uint32), ("n_site_slots", C.c_uint32), ("is_nested", C.c_uint32),
                ("n_alleles", C.c_uint64), ("n_per_base", C.c_uint64), ("n_kmer_states", C.c_uint64),
                ("device_bytes", C.c_uint64)]


@dataclass
class QuasimapReadsStats:
    """quasimap.hpp:17-24"""
    all_reads_count: int = 0
    skipped_reads_count: int = 0
    missing_kmer_reads_count: int = 0
    no_extension_reads_count: int = 0
    exact_mapped_reads_count: int = 0


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_ENC = np.zeros(256, dtype=np.uint8)
for _c, _v in zip("ACGTacgt", [1, 2, 3, 4, 1, 2, 3, 4]):
    _ENC[ord(_c)] = _v


def encode_reads(reads):
    """encode_dna_bases (src/common/utils.cpp:72-81): 1..4 = A,C,G,T; a read holding any other
    character becomes EMPTY (and is counted as skipped). Returns (bases uint8, offsets uint64)."""
    chunks, offs, t = [], [0], 0
    for r in reads:
        if isinstance(r, str):
            r = r.encode()
        e = _ENC[np.frombuffer(r, dtype=np.uint8)] if len(r) else np.zeros(0, np.uint8)
        if e.size and e.min() == 0:
            e = np.zeros(0, np.uint8)
        chunks.append(e)
        t += e.size
        offs.append(t)
    bases = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, dtype=np.uint8), np.asarray(offs, dtype=np.uint64)


class QuasimapIndex:
    """One PRG index resident on one GPU + its coverage accumulators."""

    def __init__(self, prg, kmer_size, device=0):
        self._lib = load_library()
        prg = np.ascontiguousarray(prg, dtype=np.uint32)
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.gq_index_build(_ptr(prg, C.c_uint32), prg.size, int(kmer_size), int(device), C.byref(h)))
        self._h = h
        lay = GqLayout()
        self._check(self._lib.gq_index_describe(self._h, C.byref(lay)))
        self.layout = lay
        self.n_reads = 0
        self._keep = None

    def _check(self, rc):
        if rc != 0:
            raise GqError(self._lib.gq_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gq_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options -------------------------------------------------------------------------------
    def set_option(self, name, value):
        self._check(self._lib.gq_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream):
        self._check(self._lib.gq_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    # -- mapping -------------------------------------------------------------------------------
    def _args(self, bases, offsets, seeds):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = offsets.size - 1
        if seeds.size != n:
            raise GqError("one seed per read expected")
        self._keep = (bases, offsets, seeds)
        self.n_reads = n
        return _ptr(bases, C.c_uint8), _ptr(offsets, C.c_uint64), n, _ptr(seeds, C.c_uint32)

    def map_batch(self, bases, offsets, seeds):
        """handle_reads_buffer (quasimap.cpp:82-118) for one batch held in HOST memory."""
        self._check(self._lib.gq_map_batch(self._h, *self._args(bases, offsets, seeds)))

    def upload(self, bases, offsets, seeds):
        self._check(self._lib.gq_batch_upload(self._h, *self._args(bases, offsets, seeds)))

    def map_resident(self):
        self._check(self._lib.gq_map_resident(self._h))

    def run_info(self):
        a = (C.c_double * 8)()
        self._check(self._lib.gq_last_run_info(self._h, a))
        return dict(launches=int(a[0]), rerun_strands=int(a[1]), search_ms=a[2], coverage_ms=a[3],
                    pool_words=int(a[4]), h2d_bytes=int(a[5]), kernels_ms=a[6], enqueue_ms=a[7])

    # -- results -------------------------------------------------------------------------------
    def batch_status(self):
        s = np.zeros(2 * self.n_reads, dtype=np.uint8)
        if self.n_reads:
            self._check(self._lib.gq_batch_status(self._h, _ptr(s, C.c_uint8)))
        return s

    def batch_states(self):
        """-> (offsets[2n+1], counts[2n], words): per-strand final SearchState records."""
        n = C.c_uint64()
        self._check(self._lib.gq_batch_states_size(self._h, C.byref(n)))
        off = np.zeros(2 * self.n_reads + 1, dtype=np.uint64)
        cnt = np.zeros(2 * self.n_reads, dtype=np.uint32)
        words = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._check(self._lib.gq_batch_states(self._h, _ptr(off, C.c_uint64), _ptr(cnt, C.c_uint32),
                                              _ptr(words, C.c_uint32)))
        return off, cnt, words[:n.value]

    def coverage(self):
        """-> (allele_sum uint16[n_alleles], per_base uint16[n_per_base], stats)"""
        a = np.zeros(max(self.layout.n_alleles, 1), dtype=np.uint16)
        p = np.zeros(max(self.layout.n_per_base, 1), dtype=np.uint16)
        s = np.zeros(5, dtype=np.uint64)
        self._check(self._lib.gq_coverage_fetch(self._h, _ptr(a, C.c_uint16), _ptr(p, C.c_uint16), _ptr(s, C.c_uint64)))
        st = QuasimapReadsStats(*[int(x) for x in s])
        return a[:self.layout.n_alleles], p[:self.layout.n_per_base], st

    def grouped(self):
        """flat records [site_slot, count, n, alleles...] sorted by (site_slot, alleles)."""
        n = C.c_uint64()
        self._check(self._lib.gq_coverage_grouped(self._h, None, C.byref(n)))
        w = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._check(self._lib.gq_coverage_grouped(self._h, _ptr(w, C.c_uint32), C.byref(n)))
        return w[:n.value]

    def allele_offsets(self):
        o = np.zeros(self.layout.n_site_slots + 1, dtype=np.uint64)
        self._check(self._lib.gq_index_allele_offsets(self._h, _ptr(o, C.c_uint64)))
        return o

    def per_base_layout(self):
        o = np.zeros(max(2 * self.layout.n_alleles, 1), dtype=np.uint64)
        self._check(self._lib.gq_index_per_base_layout(self._h, _ptr(o, C.c_uint64)))
        return o[:2 * self.layout.n_alleles].reshape(-1, 2)

    def read_depth_stats(self):
        """ReadStats::compute_coverage_depth -> dict(mean, variance, num_sites_noCov, num_sites_total)"""
        d = (C.c_double * 2)()
        c = np.zeros(2, dtype=np.uint64)
        self._check(self._lib.gq_read_depth_stats(self._h, d, _ptr(c, C.c_uint64)))
        return dict(mean=d[0], variance=d[1], num_sites_noCov=int(c[0]), num_sites_total=int(c[1]))

    def reset_coverage(self):
        self._check(self._lib.gq_coverage_reset(self._h))

    # -- multi-GPU plumbing ----------------------------------------------------------------------
    def device_counters(self):
        """(ptr, n_uint32, stats_ptr) of the additive accumulators, for one NCCL all-reduce."""
        p, s, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self._lib.gq_coverage_device_ptrs(self._h, C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value

    def groups_export(self):
        n = C.c_uint64()
        self._check(self._lib.gq_coverage_groups_export(self._h, None, C.byref(n)))
        w = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._check(self._lib.gq_coverage_groups_export(self._h, _ptr(w, C.c_uint32), C.byref(n)))
        return w[:n.value]

    def groups_import(self, words, replace=False):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        self._check(self._lib.gq_coverage_groups_import(self._h, _ptr(words, C.c_uint32) if words.size else None,
                                                        words.size, int(bool(replace))))
