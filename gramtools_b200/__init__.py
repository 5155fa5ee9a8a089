"""gramtools_b200 — B200-native quasimap back-end (host mirror of the C ABI in include/gq.h).

The product is ``libgq.so`` (hand-written sm_100a CUDA + C++ host code, ``gramtools_b200/csrc``).
This module is only the thin ctypes binding used by the tests, ``bench.py`` and the multi-GPU
driver; names follow the reference's quasimap interface
(libgramtools/include/genotype/quasimap/quasimap.hpp:17-32).

There is no CPU fallback: importing works anywhere, but every quasimap call needs the built
library and a CUDA device and raises otherwise. The genotyping step that follows quasimap
(``level_genotype*``) is host code in the reference too and needs no device.
"""
from .engine import (GqError, QuasimapIndex, QuasimapReadsStats, comm_unique_id, encode_reads, level_genotype,  # noqa: F401
                     level_genotype_json, lib_path, load_library, pack_ascii, pack_reads, read_depth_stats_host,
                     suffix_array)
from .synth import (make_snp_prg, make_indel_prg, make_nested_prg, sample_reads, master_seeds)  # noqa: F401

__all__ = [
    "GqError", "QuasimapIndex", "QuasimapReadsStats", "comm_unique_id", "encode_reads", "lib_path", "load_library",
    "pack_ascii", "pack_reads", "suffix_array", "level_genotype", "level_genotype_json", "read_depth_stats_host",
    "make_snp_prg", "make_indel_prg", "make_nested_prg", "sample_reads", "master_seeds",
]
