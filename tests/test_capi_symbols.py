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
