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
        "gq_suffix_array": [u32p, C.c_uint64, C.c_int, u32p, C.POINTER(C.c_int)],
        "gq_kmer_index_dump": [vp, C.c_char_p],
        "gq_index_save": [vp, C.c_char_p],
        "gq_index_load": [C.c_char_p, C.c_int, C.POINTER(vp)],
        "gq_index_prg": [vp, u32p, u64p],
        "gq_index_build_from_gram_dir": [u32p, C.c_uint64, C.c_uint32, C.c_int, C.c_char_p, C.POINTER(vp)],
        "gq_index_describe": [vp, C.POINTER(GqLayout)],
        "gq_index_allele_offsets": [vp, u64p],
        "gq_index_per_base_layout": [vp, u64p],
        "gq_map_batch": [vp, u8p, u64p, C.c_uint64, u32p],
        "gq_map_batch_packed": [vp, u32p, u32p, u32p, C.c_uint64, u32p],
        "gq_packed_words": [u64p, C.c_uint64, u64p],
        "gq_pack_reads": [u8p, u64p, C.c_uint64, u32p, u32p, u32p, C.c_int],
        "gq_pack_ascii": [C.c_char_p, u64p, C.c_uint64, u32p, u32p, u32p, C.c_int],
        "gq_host_alloc": [C.c_uint64, C.POINTER(vp)],
        "gq_host_free": [vp],
        "gq_device_count": [C.POINTER(C.c_int)],
        "gq_comm_unique_id": [u8p],
        "gq_comm_init": [vp, u8p, C.c_int, C.c_int],
        "gq_comm_init_all": [C.POINTER(vp), C.c_int],
        "gq_comm_destroy": [vp],
        "gq_comm_version": [C.POINTER(C.c_int)],
        "gq_coverage_allreduce": [vp],
        "gq_coverage_allreduce_all": [C.POINTER(vp), C.c_int],
        "gq_index_clone": [vp, C.c_int, C.POINTER(vp)],
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
        "gq_last_kernel_ms": [vp, C.POINTER(C.c_double)],
        "gq_level_genotype": [u32p, C.c_uint64, u16p, C.c_uint64, u32p, C.c_uint64, C.POINTER(C.c_double), C.c_int,
                              C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_int],
        "gq_level_genotype_json": [u32p, C.c_uint64, u16p, C.c_uint64, u32p, C.c_uint64, C.POINTER(C.c_double), C.c_int,
                                   C.c_char_p, C.c_uint32, C.c_int, C.c_char_p, u64p],
        "gq_read_depth_stats_host": [u32p, C.c_uint64, u16p, C.c_uint64, u32p, C.c_uint64, C.POINTER(C.c_double), u64p],
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


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through libgq (rank 0 calls it and hands the 128 bytes to the other ranks)."""
    lib = load_library()
    a = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
    if lib.gq_comm_unique_id(_ptr(a, C.c_uint8)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return a


def suffix_array(prg, device=0):
    """gq_suffix_array: SA of the PRG (+ sentinel) built on the GPU by prefix doubling -> (sa uint32[n + 1], rounds)."""
    lib = load_library()
    prg = np.ascontiguousarray(prg, dtype=np.uint32)
    sa = np.zeros(prg.size + 1, dtype=np.uint32)
    rounds = C.c_int(0)
    if lib.gq_suffix_array(_ptr(prg, C.c_uint32), prg.size, device, _ptr(sa, C.c_uint32), C.byref(rounds)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return sa, rounds.value


def pack_reads(bases, offsets, n_threads=None, out=None):
    """gq_pack_reads: encoded bases (1..4) -> (packed uint32, word_off uint32[n+1], len uint32[n]), the host form
    gq_map_batch_packed takes. `out` = optional preallocated (packed, word_off, len), e.g. pinned memory."""
    lib = load_library()
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    nw = C.c_uint64()
    if lib.gq_packed_words(_ptr(offsets, C.c_uint64), n, C.byref(nw)) != 0:
        raise GqError(lib.gq_last_error().decode())
    if out is None:
        out = (np.zeros(nw.value, np.uint32), np.zeros(n + 1, np.uint32), np.zeros(max(n, 1), np.uint32))
    packed, word_off, ln = out
    assert packed.size >= nw.value and word_off.size >= n + 1 and ln.size >= n
    if lib.gq_pack_reads(_ptr(bases, C.c_uint8), _ptr(offsets, C.c_uint64), n, _ptr(packed, C.c_uint32),
                         _ptr(word_off, C.c_uint32), _ptr(ln, C.c_uint32), int(n_threads or os.cpu_count() or 1)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return packed, word_off, ln[:n]


def pack_ascii(text, offsets, n_threads=None):
    """gq_pack_ascii: sequence text (bytes) -> packed form; reads with a non-ACGT character become empty."""
    lib = load_library()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    nw = C.c_uint64()
    lib.gq_packed_words(_ptr(offsets, C.c_uint64), n, C.byref(nw))
    packed, word_off, ln = np.zeros(nw.value, np.uint32), np.zeros(n + 1, np.uint32), np.zeros(max(n, 1), np.uint32)
    if lib.gq_pack_ascii(bytes(text), _ptr(offsets, C.c_uint64), n, _ptr(packed, C.c_uint32), _ptr(word_off, C.c_uint32),
                         _ptr(ln, C.c_uint32), int(n_threads or os.cpu_count() or 1)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return packed, word_off, ln[:n]


class QuasimapIndex:
    """One PRG index resident on one GPU + its coverage accumulators."""

    def __init__(self, prg, kmer_size, device=0, _clone_of=None, kmer_index_dir=None, index_file=None):
        """kmer_index_dir: take the k-mer index from the sdsl files of that gram_dir (gq_index_build_from_gram_dir);
        index_file: load a whole index written by save() (gq_index_load; prg / kmer_size are ignored)."""
        self._lib = load_library()
        h = C.c_void_p()
        self._h = None
        if index_file is not None:
            self._check(self._lib.gq_index_load(os.fsencode(index_file), int(device), C.byref(h)))
        elif _clone_of is not None:
            self._check(self._lib.gq_index_clone(_clone_of._h, int(device), C.byref(h)))
        elif kmer_index_dir is not None:
            prg = np.ascontiguousarray(prg, dtype=np.uint32)
            self._check(self._lib.gq_index_build_from_gram_dir(_ptr(prg, C.c_uint32), prg.size, int(kmer_size), int(device),
                                                              os.fsencode(kmer_index_dir), C.byref(h)))
        else:
            prg = np.ascontiguousarray(prg, dtype=np.uint32)
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

    def save(self, path):
        """gq_index_save: the whole flat index as one checksummed file."""
        self._check(self._lib.gq_index_save(self._h, os.fsencode(path)))

    def prg(self):
        n = C.c_uint64()
        self._check(self._lib.gq_index_prg(self._h, None, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._check(self._lib.gq_index_prg(self._h, _ptr(out, C.c_uint32), C.byref(n)))
        return out[:n.value]

    def kmer_index_dump(self, gram_dir):
        """Write kmers / kmers_stats / sa_intervals / paths (the reference's gram_dir files of the k-mer index)."""
        self._check(self._lib.gq_kmer_index_dump(self._h, os.fsencode(gram_dir)))

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

    def map_batch_packed(self, packed, word_off, length, seeds):
        """The same batch from 2-bit packed HOST buffers (pack_reads / pack_ascii)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        word_off = np.ascontiguousarray(word_off, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = word_off.size - 1
        if seeds.size != n or length.size < n:
            raise GqError("one seed and one length per read expected")
        self._keep = (packed, word_off, length, seeds)
        self.n_reads = n
        self._check(self._lib.gq_map_batch_packed(self._h, _ptr(packed, C.c_uint32), _ptr(word_off, C.c_uint32),
                                                  _ptr(length, C.c_uint32), n, _ptr(seeds, C.c_uint32)))

    def clone(self, device):
        """A second handle on another GPU of this process (index built once on the host)."""
        return QuasimapIndex(None, 0, device=device, _clone_of=self)

    # -- multi-GPU: one exchange at the end (include/gq.h) ------------------------------------------
    def comm_init(self, comm_id, rank, n_ranks):
        comm_id = np.ascontiguousarray(comm_id, dtype=np.uint8)
        assert comm_id.size == COMM_ID_BYTES
        self._check(self._lib.gq_comm_init(self._h, _ptr(comm_id, C.c_uint8), int(rank), int(n_ranks)))

    def coverage_allreduce(self):
        """In-place sum of the coverage of all ranks (dense counters, stats, sparse groups) over NCCL."""
        self._check(self._lib.gq_coverage_allreduce(self._h))

    @staticmethod
    def comm_init_all(handles):
        lib = load_library()
        arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
        if lib.gq_comm_init_all(arr, len(handles)) != 0:
            raise GqError(lib.gq_last_error().decode())

    @staticmethod
    def coverage_allreduce_all(handles):
        lib = load_library()
        arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
        if lib.gq_coverage_allreduce_all(arr, len(handles)) != 0:
            raise GqError(lib.gq_last_error().decode())

    def upload(self, bases, offsets, seeds):
        self._check(self._lib.gq_batch_upload(self._h, *self._args(bases, offsets, seeds)))

    def map_resident(self):
        self._check(self._lib.gq_map_resident(self._h))

    def run_info(self):
        a = (C.c_double * 8)()
        self._check(self._lib.gq_last_run_info(self._h, a))
        return dict(launches=int(a[0]), rerun_strands=int(a[1]), search_ms=a[2], coverage_ms=a[3],
                    pool_words=int(a[4]), h2d_bytes=int(a[5]), kernels_ms=a[6], enqueue_ms=a[7])

    KERNELS = ["seed_kernel", "verify_kernel", "text_kernel", "search_kernel", "classify_kernel", "coverage_kernel",
               "revcomp_kernel"]

    def kernel_ms(self):
        """Per-kernel durations of the last single-slice map_resident (CUDA events inside the library)."""
        a = (C.c_double * 8)()
        self._check(self._lib.gq_last_kernel_ms(self._h, a))
        return dict(zip(self.KERNELS, [float(x) for x in a[:7]]))

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


# ---- genotyping step (host code of libgq.so; genotype.cpp:68-118) ----------------------------------------------

def _cov_args(prg, per_base, grouped):
    prg = np.ascontiguousarray(prg, dtype=np.uint32)
    per_base = np.ascontiguousarray(per_base, dtype=np.uint16)
    grouped = np.ascontiguousarray(grouped, dtype=np.uint32)
    pb = per_base if per_base.size else np.zeros(1, np.uint16)
    gw = grouped if grouped.size else np.zeros(1, np.uint32)
    keep = (prg, pb, gw)
    return keep, (_ptr(prg, C.c_uint32), prg.size, _ptr(pb, C.c_uint16), per_base.size, _ptr(gw, C.c_uint32), grouped.size)


def read_depth_stats_host(prg, per_base, grouped):
    """ReadStats::compute_coverage_depth (read_stats.cpp:119-160) from fetched coverage, without a device handle."""
    lib = load_library()
    keep, args = _cov_args(prg, per_base, grouped)
    d = (C.c_double * 2)()
    c = np.zeros(2, dtype=np.uint64)
    if lib.gq_read_depth_stats_host(*args, d, _ptr(c, C.c_uint64)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return {"mean": d[0], "variance": d[1], "num_sites_noCov": int(c[0]), "num_sites_total": int(c[1])}


def level_genotype_json(prg, per_base, grouped, mean_cov, var_cov, mean_pb_error, ploidy="haploid", sample_id="sample",
                        gcp_seed=42, n_threads=1):
    """LevelGenotyper + make_json_prg (runner.cpp:27-97, make_json.cpp:7-25): the text of genotyped.json."""
    lib = load_library()
    keep, args = _cov_args(prg, per_base, grouped)
    stats = (C.c_double * 3)(mean_cov, var_cov, mean_pb_error)
    pl = {"haploid": 1, "diploid": 2}[ploidy]
    n = np.zeros(1, dtype=np.uint64)
    if lib.gq_level_genotype_json(*args, stats, pl, sample_id.encode(), gcp_seed, n_threads, None, _ptr(n, C.c_uint64)) != 0:
        raise GqError(lib.gq_last_error().decode())
    buf = C.create_string_buffer(int(n[0]))
    if lib.gq_level_genotype_json(*args, stats, pl, sample_id.encode(), gcp_seed, n_threads, buf, _ptr(n, C.c_uint64)) != 0:
        raise GqError(lib.gq_last_error().decode())
    return buf.value.decode()


def level_genotype(prg, per_base, grouped, mean_cov, var_cov, mean_pb_error, genotype_dir, ploidy="haploid",
                   sample_id="sample", prg_coords_path=None, debug_path=None, gcp_seed=42, n_threads=1):
    """The genotyping half of commands::genotype::run (genotype.cpp:68-118): writes genotyped.json,
    personalised_reference.fasta and genotyped.vcf.gz into `genotype_dir`."""
    lib = load_library()
    keep, args = _cov_args(prg, per_base, grouped)
    stats = (C.c_double * 3)(mean_cov, var_cov, mean_pb_error)
    pl = {"haploid": 1, "diploid": 2}[ploidy]
    rc = lib.gq_level_genotype(*args, stats, pl, sample_id.encode(),
                               os.fsencode(prg_coords_path) if prg_coords_path else None, os.fsencode(genotype_dir),
                               os.fsencode(debug_path) if debug_path else None, gcp_seed, n_threads)
    if rc != 0:
        raise GqError(lib.gq_last_error().decode())
