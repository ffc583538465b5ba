"""The C-ABI library loads and exports every symbol include/orcdemux.h declares (no compute
calls here: this runs without a GPU), and refuses to work without a device."""
import ctypes as C
import os
import re

import pytest

import helpers as H
from orcdemux import lib

HEADER = os.path.join(H.ROOT, "include", "orcdemux.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if not os.path.exists(lib.SO_PATH):
        import __graft_entry__ as g
        g.build()
    L = lib.load()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), "liborcdemux.so does not export %s" % n
    assert sorted(lib.EXPORTS) == names
    assert b"sm_100a" in L.orc_version()


def test_struct_sizes_match_header():
    # spot-check the layouts ctypes mirrors
    assert C.sizeof(lib.RoundParams) == 48
    assert C.sizeof(lib.Batch) == 88
    assert lib.MATCH_DTYPE.itemsize == 32


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from orcdemux import engine as E
    with pytest.raises(E.OrcError, match="CUDA"):
        E.Engine(E.m13_rounds(), max_reads=16, max_bytes=1024)


def test_product_does_not_import_oracle():
    pkg = os.path.join(H.PKG)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt, f
