"""libgq.so loads without a GPU and exports every entry point include/gq.h declares; compute entry
points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "gq.h")).read()
    return sorted(set(re.findall(r"\b(gq_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_header_symbols(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libgq.so does not export {n}"


def test_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gramtools_b200 import GqError, QuasimapIndex
    with pytest.raises(GqError, match="no CUDA device|CUDA"):
        QuasimapIndex(np.asarray([1, 2, 3, 4], dtype=np.uint32), 2)


def test_product_does_not_link_the_oracle(built_lib):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    assert "gqo_" not in out and "emu_" not in out
    srcs = os.path.join(ROOT, "gramtools_b200")
    for root, _, files in os.walk(srcs):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".py", "Makefile")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle/" not in txt and "gq_oracle" not in txt, f


def test_host_packers(built_lib):
    """gq_pack_reads / gq_pack_ascii (host code, no GPU): 16 bases per word, base j at bits [2j, 2j+2), read r at
    word (offset >> 4) + r; text with a non-ACGT character becomes an empty read (utils.cpp:13-47,83-92)."""
    from gramtools_b200 import encode_reads, pack_ascii, pack_reads
    rng = np.random.default_rng(5)
    reads = ["".join("ACGT"[x] for x in rng.integers(0, 4, int(L))) for L in rng.integers(0, 70, 200)]
    reads[3] = "ACGTNNACGT"
    reads[7] = "acgtacgtacgtacgtacgt"
    bases, offs = encode_reads(reads)
    packed, word_off, ln = pack_reads(bases, offs, n_threads=3)
    for r in range(len(reads)):
        L = int(offs[r + 1] - offs[r])
        assert ln[r] == L and word_off[r] == (int(offs[r]) >> 4) + r
        for j in range(L):
            assert (int(packed[word_off[r] + j // 16]) >> (2 * (j % 16))) & 3 == int(bases[int(offs[r]) + j]) - 1
    assert word_off[len(reads)] >= word_off[len(reads) - 1] + (int(ln[-1]) + 15) // 16
    # the same reads from text: identical words for clean reads, length 0 for the read with Ns
    text = "".join(reads).encode()
    toff = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=toff[1:])
    p2, w2, l2 = pack_ascii(text, toff, n_threads=2)
    assert l2[3] == 0 and l2[7] == 20
    for r in range(len(reads)):
        if r == 3:
            continue
        assert l2[r] == len(reads[r])
        for j in range(len(reads[r])):
            assert (int(p2[w2[r] + j // 16]) >> (2 * (j % 16))) & 3 == "ACGT".index(reads[r][j].upper())


def test_plain_c_client_compiles_links_and_runs(built_lib, tmp_path):
    """include/gq.h is a C header (C99, no C++), libgq.so links from C, host entry points work without a GPU."""
    import subprocess
    exe = str(tmp_path / "client")
    lib_dir = os.path.dirname(built_lib)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_client", "client.c"), "-o", exe, "-L", lib_dir, "-lgq", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "c client ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
