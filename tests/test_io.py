"""Host-side FASTQ(.gz) streaming of liborcdemux.so (csrc/orc_io.cpp: orc_reader_* / orc_writer_*),
the plumbing dnaio + xopen + the ordered chunk writer provide around cutadapt's matching.  Bytes
in, bytes out: no GPU and no matching involved, so this runs on CPU."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from orcdemux import fastq as F
from orcdemux import synth


class _Res:
    """What BinWriters needs of an engine.BatchResult."""

    def __init__(self, parts):
        self.fastq = np.frombuffer(b"".join(parts), dtype=np.uint8)
        self.bin_offsets = np.cumsum([0] + [len(p) for p in parts]).astype(np.uint64)


@pytest.fixture(scope="module")
def reads():
    rs = synth.generate(6000, 300, 900, seed=5)
    return rs, rs.to_fastq_bytes()


@pytest.mark.parametrize("gz", [True, False])
def test_reader_batches_cover_the_file(tmp_path, reads, gz):
    rs, raw = reads
    p = tmp_path / ("in.fastq.gz" if gz else "in.fastq")
    if gz:
        # two gzip members: a concatenation is one stream (what the writer below produces)
        with open(p, "wb") as fh:
            fh.write(gzip.compress(raw[:len(raw) // 3], 1) + gzip.compress(raw[len(raw) // 3:], 1))
    else:
        p.write_bytes(raw)
    with F.FastqReader(str(p), max_reads=1000, max_bytes=1 << 20, keep=2, ahead=2, pinned=False) as rd:
        n, chunks, prev = 0, [], None
        for tb in rd:
            assert tb.n_reads <= 1000 and tb.n_bytes <= 1 << 20
            if prev is not None:                    # the previous batch is still valid (keep == 2)
                assert prev[0].read(0) == prev[1]
            prev = (tb, tb.read(0))
            assert tb.read(tb.n_reads - 1) == rs.read(n + tb.n_reads - 1)
            assert tb.total_bases() == int(tb.lengths.sum())
            n += tb.n_reads
            chunks.append(tb.text[:tb.n_bytes].tobytes())
    assert n == rs.n_reads and b"".join(chunks) == raw


def test_reader_stdin_and_last_record_without_newline(tmp_path, reads):
    _, raw = reads
    p = tmp_path / "in.fastq"
    p.write_bytes(raw[:-1])
    code = ("import sys; sys.path.insert(0, %r); from orcdemux import fastq as F\n"
            "print(sum(tb.n_reads for tb in F.FastqReader('-', 4096, 1 << 22, pinned=False)))" % H.PKG)
    with open(p, "rb") as fh:
        out = subprocess.run([sys.executable, "-c", code], stdin=fh, capture_output=True, text=True, check=True)
    assert int(out.stdout.strip()) == 6000


def test_reader_errors(tmp_path):
    with pytest.raises(OSError, match="cannot open"):
        F.FastqReader(str(tmp_path / "missing.fastq.gz"), pinned=False)
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\nIII\n")
    with pytest.raises(ValueError, match="differ in length"):
        list(F.FastqReader(str(bad), 16, 1 << 12, pinned=False))
    big = tmp_path / "big.fastq"
    big.write_bytes(b"@r\n" + b"A" * 5000 + b"\n+\n" + b"I" * 5000 + b"\n")
    with pytest.raises(ValueError, match="larger than the batch buffer"):
        list(F.FastqReader(str(big), 16, 4096, pinned=False))
    trunc = tmp_path / "trunc.fastq.gz"
    trunc.write_bytes(gzip.compress(b"@r\nACGT\n+\nIIII\n" * 1000)[:-20])
    with pytest.raises(ValueError, match="reading the input"):
        list(F.FastqReader(str(trunc), 4096, 1 << 16, pinned=False))
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    assert list(F.FastqReader(str(empty), 16, 4096, pinned=False)) == []


def test_writer_order_members_and_empty_bins(tmp_path, reads):
    _, raw = reads
    paths = [str(tmp_path / "b0.fastq.gz"), None, str(tmp_path / "b2.fastq"), str(tmp_path / "b3.fastq.gz")]
    w = F.BinWriters(paths, 5, threads=4)
    assert all(os.path.exists(p) for p in paths if p)          # created up front (02:75-85 lists them)
    exp = [b"", b"", b"", b""]
    rng = np.random.default_rng(1)
    tickets = []
    for it in range(8):
        parts = [raw[int(rng.integers(0, 1000)):int(rng.integers(1000, len(raw)))] for _ in range(4)]
        parts[3] = b""
        if it % 2:
            parts[0] = b""
        for b in range(4):
            if paths[b]:
                exp[b] += parts[b]
        tickets.append(w.write_batch(_Res(parts)))
        if it >= 2:
            w.wait(tickets[it - 2])
    w.close()
    w.close()                                                   # idempotent
    assert gzip.open(paths[0]).read() == exp[0]                 # > 4 MiB per batch: several members each
    assert open(paths[2], "rb").read() == exp[2]
    assert os.path.getsize(paths[3]) > 0 and gzip.open(paths[3]).read() == b""
    assert w.bytes_written == [len(exp[0]), 0, len(exp[2]), 0]


def test_writer_errors(tmp_path):
    with pytest.raises(OSError, match="cannot create"):
        F.BinWriters([str(tmp_path / "no" / "such" / "dir.fastq.gz")], 5, 2)


def test_truncated_header_is_an_error():
    """text cut off in the middle of a header line ("...IIII\\n@b", no newline) used to index one
    read and report success; dnaio raises on such input."""
    import ctypes as C
    from orcdemux import lib as LIB
    L = LIB.load()
    for tail, ok in ((b"", True), (b"\n\n", True), (b"@b", False), (b"@b\nAC", False), (b"@b\nAC\n+\n", False)):
        text = np.frombuffer(b"@a\nACGT\n+\nIIII\n" + tail, dtype=np.uint8).copy()
        so, qo, no = (np.zeros(4, np.uint64) for _ in range(3))
        ln, nl = np.zeros(4, np.uint32), np.zeros(4, np.uint32)
        used = C.c_uint64(0)
        err = C.create_string_buffer(256)
        n = L.orc_fastq_index(text.ctypes.data, text.shape[0], 4, 1, so.ctypes.data, ln.ctypes.data, qo.ctypes.data,
                              no.ctypes.data, nl.ctypes.data, C.byref(used), err, 256)
        if ok:
            assert n == 1 and used.value == text.shape[0], (tail, n, err.value)
        else:
            assert n == LIB.ORC_EINVAL and (b"truncated" in err.value or b"differ in length" in err.value), (tail, n, err.value)
        # not final: the partial record is simply left for the next call
        n = L.orc_fastq_index(text.ctypes.data, text.shape[0], 4, 0, so.ctypes.data, ln.ctypes.data, qo.ctypes.data,
                              no.ctypes.data, nl.ctypes.data, C.byref(used), err, 256)
        assert n == 1 and used.value == 15


def test_member_parallel_inflate_and_index_sidecar(tmp_path, reads):
    """Files orc_writer wrote carry a size field per gzip member: orc_reader inflates them on several threads and
    returns the same text as one zlib stream would; python's gzip reads them too; PATH.idx lists the chunks;
    a foreign member in the middle hands the rest of the file to the serial reader; a cut file is an error."""
    rs, raw = reads
    cut = [0, len(raw) // 5, len(raw) // 2, 3 * len(raw) // 4, len(raw)]
    # cut at record boundaries so that every batch holds whole records
    cut = [0] + [raw.index(b"\n@r", c) + 1 for c in cut[1:-1]] + [len(raw)]
    paths = [str(tmp_path / "a.fastq.gz"), str(tmp_path / "empty.fastq.gz")]
    w = F.BinWriters(paths, 1, threads=3, index=True)
    w._L.orc_writer_set_index(w._w, 1)
    tickets = [w.write_batch(_Res([raw[a:b], b""])) for a, b in zip(cut[:-1], cut[1:])]
    for t in tickets:
        w.wait(t)
    w.close()
    blob = open(paths[0], "rb").read()
    assert gzip.decompress(blob) == raw and gzip.decompress(open(paths[1], "rb").read()) == b""
    idx = np.fromfile(paths[0] + ".idx", dtype="<u8").reshape(-1, 2)
    assert int(idx[:, 1].sum()) == len(blob) and list(idx[:, 0]) == sorted(idx[:, 0]) and set(idx[:, 0]) == set(tickets)
    assert np.fromfile(paths[1] + ".idx", dtype="<u8").size == 0
    # the size field: FEXTRA, subfield OC, the member's own length
    assert blob[3] & 4 and blob[12:14] == b"OC" and int.from_bytes(blob[16:20], "little") == int(idx[0, 1])

    def text_of(path, threads):
        with F.FastqReader(path, max_reads=700, max_bytes=1 << 20, keep=2, ahead=2, pinned=False, threads=threads) as rd:
            return b"".join(tb.text[:tb.n_bytes].tobytes() for tb in rd)

    for threads in (1, 4):
        assert text_of(paths[0], threads) == raw
    assert text_of(paths[1], 4) == b""
    # our members, then two foreign ones, then ours again: everything after the first foreign member is read serially
    mixed = tmp_path / "mixed.fastq.gz"
    k = int(idx[0, 1])
    a, b = cut[1], cut[2]
    mixed.write_bytes(blob[:k] + gzip.compress(raw[a:b], 1) + gzip.compress(raw[b:cut[3]], 6))
    first = gzip.decompress(blob[:k])
    assert text_of(str(mixed), 4) == first + raw[a:cut[3]]
    trunc = tmp_path / "trunc.fastq.gz"
    trunc.write_bytes(blob[:len(blob) - 9])
    with pytest.raises(ValueError, match="input"):
        text_of(str(trunc), 4)
