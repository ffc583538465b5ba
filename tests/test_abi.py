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


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "SO_PATH", str(tmp_path / "liborcdemux.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        lib.load()


def test_cli_surface_parsing():
    from orcdemux import cli
    opt = cli.parse_cutadapt_argv("--action=trim -e 0.1 -j 24 --rc -g file:x.fa -o o/{name}_d.fastq.gz in.fastq.gz "
                                  "--json=o/r.json".split())
    assert opt["rc"] and opt["e"] == 0.1 and opt["cores"] == 24 and opt["json"] == "o/r.json"
    assert opt["g"] == ["file:x.fa"] and opt["inputs"] == ["in.fastq.gz"]
    opt = cli.parse_cutadapt_argv("-a file:y.fa -O 5 --no-indels -e0.2 -o {name}.fq in.fq".split())
    assert opt["O"] == 5 and not opt["indels"] and opt["e"] == 0.2 and opt["a"] == ["file:y.fa"]
    for bad in ("-m 10 -g AAA -o {name}.fq in.fq", "-g AAA -a CCC -o {name} in", "-g AAA --discard-untrimmed -o {name}.fq in", "-g AAA in",
                "--action=mask -g AAA -o {name} in", "-g AAA -o {name} a.fq b.fq"):
        with pytest.raises(cli.Unsupported):
            cli.parse_cutadapt_argv(bad.split())
    assert cli.parse_cutadapt_argv("-g AAA -o out.fq in".split())["out"] == "out.fq"      # one output: primers.py
    assert cli.dataset_name("/x/pychopped/pychopped_s1_pass.fastq.gz") == "s1"
    names, seqs, anchored = cli._parse_adapter_specs(["first=ACGT", "TTGCA"], 0)
    assert names == ["first", "2"] and seqs == ["ACGT", "TTGCA"] and not anchored
    assert cli._parse_adapter_specs(["^ACGT"], 0)[2] and cli._parse_adapter_specs(["ACGT$"], 1)[2]


def test_fastq_indexer_edge_cases():
    import numpy as np
    from orcdemux import fastq as F
    txt = b"@r1 c\nACGT\n+\nIIII\n@r2\nAC\r\n+r2\nII\r\n@r3\n\n+\n\n@r4\nAAA\n+\nIII"
    a = np.frombuffer(txt, dtype=np.uint8).copy()
    n, used, arr = F.index_text(a, len(a), 10, False)
    assert n == 3 and used == 42
    n, used, arr = F.index_text(a, len(a), 10, True)          # final: last record without newline
    assert n == 4 and used == len(a) and list(arr[1][:4]) == [4, 2, 0, 3]
    tb = F.TextBatch(a, used, n, *arr)
    assert tb.read(1) == ("r2", "AC", "II") and tb.read(2) == ("r3", "", "") and tb.read(3) == ("r4", "AAA", "III")
    bad = np.frombuffer(b"@r\nACGT\n+\nII\n", dtype=np.uint8).copy()
    with pytest.raises(ValueError, match="differ in length"):
        F.index_text(bad, len(bad), 10, True)
    bad = np.frombuffer(b"r\nACGT\n+\nIIII\n", dtype=np.uint8).copy()
    with pytest.raises(ValueError, match="'@'"):
        F.index_text(bad, len(bad), 10, True)
