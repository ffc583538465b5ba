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


def _text_and_mode(path, threads, max_bytes=1 << 20):
    with F.FastqReader(str(path), max_reads=4000, max_bytes=max_bytes, keep=2, ahead=2, pinned=False, threads=threads) as rd:
        chunks = [tb.text[:tb.n_bytes].tobytes() for tb in rd]
        return b"".join(chunks), rd.inflate_mode()


@pytest.fixture(scope="module")
def big_text():
    """About 19 MB of FASTQ with qualities that do not repeat (gzip leaves 8-9 MB of it)."""
    rs = synth.generate(14000, 300, 900, seed=11)
    raw = bytearray(rs.to_fastq_bytes())
    rng = np.random.default_rng(3)
    off = 0
    for i in range(rs.n_reads):            # the qualities line of every record: a random walk like a basecaller's
        name, seq, _ = rs.read(i)
        q0 = off + 1 + len(name) + 1 + len(seq) + 3
        steps = rng.integers(-2, 3, len(seq))
        raw[q0:q0 + len(seq)] = (33 + np.clip(20 + np.cumsum(steps), 2, 50)).astype(np.uint8).tobytes()
        off = q0 + len(seq) + 1
    assert off == len(raw)
    return bytes(raw)


def test_chunk_parallel_inflate_of_a_foreign_gzip(tmp_path, big_text, monkeypatch):
    """A .gz this library did not write (02_cutadapt_loop.sh:64-72 reads pychopped_<dataset>.fastq.gz) is inflated
    chunk by chunk on the pool (csrc/orc_pgz.h) and gives the bytes one zlib stream gives: levels 1-9, the
    `gzip` program, chunks far smaller than a DEFLATE block (so that most block starts have to be bridged by the
    reader thread), one thread per chunk; small files, stdin and ORC_NO_PGZ keep the zlib stream."""
    raw = big_text
    monkeypatch.setenv("ORC_PGZ_MIN", "1000000")
    p = tmp_path / "in.fastq.gz"
    for level in (1, 6, 9):
        p.write_bytes(gzip.compress(raw, level))
        for chunk, threads in ((1 << 20, 4), (1 << 18, 3), (30000, 8)):
            monkeypatch.setenv("ORC_PGZ_CHUNK", str(chunk))
            text, (mode, par, ser) = _text_and_mode(p, threads)
            assert mode == 2 and par + ser == len(raw), (level, chunk, mode, par, ser)
            assert text == raw, (level, chunk)
            if chunk >= 1 << 18:
                assert par > 0.9 * len(raw), (level, chunk, par, ser)      # the pool did the work
    monkeypatch.setenv("ORC_PGZ_CHUNK", str(1 << 20))
    plain = tmp_path / "x.fastq"
    plain.write_bytes(raw)
    subprocess.run(["gzip", "-k", "-f", str(plain)], check=True)
    text, (mode, par, ser) = _text_and_mode(str(plain) + ".gz", 4, max_bytes=1 << 22)
    assert mode == 2 and text == raw and ser == 0
    # one inflate thread, the switch, a small file: one zlib stream
    assert _text_and_mode(p, 1)[1][0] == 0
    monkeypatch.setenv("ORC_NO_PGZ", "1")
    text, (mode, _, _) = _text_and_mode(p, 4)
    assert mode == 0 and text == raw
    monkeypatch.delenv("ORC_NO_PGZ")
    small = tmp_path / "small.fastq.gz"
    small.write_bytes(gzip.compress(raw[:raw.index(b"\n@r", 200000) + 1], 6))
    assert _text_and_mode(small, 4)[1][0] == 0


def test_chunk_parallel_inflate_of_other_streams(tmp_path, big_text, monkeypatch):
    """Members that end inside a chunk (cat a.gz b.gz ..., the usual way MinKNOW's files are joined), a header with
    every optional field, stored and fixed-code blocks, full-flush points, bytes behind the last member."""
    import zlib
    raw = big_text
    monkeypatch.setenv("ORC_PGZ_MIN", "100000")
    monkeypatch.setenv("ORC_PGZ_CHUNK", str(1 << 19))
    cuts = [0] + [raw.index(b"\n@r", c) + 1 for c in range(700000, len(raw) - 700000, 700000)] + [len(raw)]
    p = tmp_path / "cat.fastq.gz"
    p.write_bytes(b"".join(gzip.compress(raw[a:b], 6) for a, b in zip(cuts[:-1], cuts[1:])))
    text, (mode, par, ser) = _text_and_mode(p, 4)
    assert mode == 2 and text == raw and par > 0.9 * len(raw), (par, ser)
    # FEXTRA + FNAME + FCOMMENT + FHCRC in front of a raw stream with sync and full flushes
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = []
    for i in range(0, len(raw), 250000):
        parts.append(co.compress(raw[i:i + 250000]))
        parts.append(co.flush(zlib.Z_FULL_FLUSH if (i // 250000) % 3 else zlib.Z_SYNC_FLUSH))
    parts.append(co.flush())
    hdr = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"ab\x01\x00z" + b"n.fq\0" + b"c\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")
    p.write_bytes(hdr + b"".join(parts) + zlib.crc32(raw).to_bytes(4, "little") + (len(raw) & 0xffffffff).to_bytes(4, "little") + b"\0" * 512)
    text, (mode, par, ser) = _text_and_mode(p, 4)
    assert mode == 2 and text == raw
    # fixed codes only, Huffman only, stored only (nothing for the pool to find: the reader thread decodes it)
    part = raw[:raw.index(b"\n@r", 3000000) + 1]
    for strategy, level in ((zlib.Z_FIXED, 6), (zlib.Z_HUFFMAN_ONLY, 6), (zlib.Z_DEFAULT_STRATEGY, 0)):
        co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
        p.write_bytes(co.compress(part) + co.flush())
        text, (mode, par, ser) = _text_and_mode(p, 3)
        assert mode == 2 and text == part, (strategy, level)


def test_chunk_parallel_inflate_refuses_damaged_input(tmp_path, big_text, monkeypatch):
    """A cut stream, a flipped bit in the codes, a wrong CRC-32 or length in the trailer: an error as from zlib, in
    whatever chunk the damage lies."""
    raw = big_text[:big_text.index(b"\n@r", 6000000) + 1]
    good = gzip.compress(raw, 6)
    monkeypatch.setenv("ORC_PGZ_MIN", "100000")
    monkeypatch.setenv("ORC_PGZ_CHUNK", str(1 << 18))
    p = tmp_path / "bad.fastq.gz"
    rng = np.random.default_rng(5)
    cases = [good[:len(good) // 2], good[:-5], good[:-8] + b"\0\0\0\0" + good[-4:], good[:-4] + b"\1\0\0\0"]
    for _ in range(6):
        b = bytearray(good)
        b[int(rng.integers(100, len(good) - 100))] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    for i, blob in enumerate(cases):
        p.write_bytes(blob)
        # ("reading the input: ..." from the inflate, or the indexer's complaint about text that a damaged code
        # turned into something else -- that batch is handed on before the stream's CRC is reached, as with gzread)
        with pytest.raises(ValueError):
            _text_and_mode(p, 4)
        with pytest.raises(Exception):
            gzip.decompress(blob)


def test_bgzf_members_are_inflated_in_parallel(tmp_path, reads):
    """bgzip / htslib output: members of at most 64 KiB with the "BC" size subfield -- hopped over and inflated on
    the pool like this library's own members."""
    import zlib
    _, raw = reads

    def block(data):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(data) + co.flush()
        total = 18 + len(body) + 8
        return (bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0]) + b"BC\x02\x00" + (total - 1).to_bytes(2, "little") + body +
                zlib.crc32(data).to_bytes(4, "little") + len(data).to_bytes(4, "little"))

    blob = b"".join(block(raw[i:i + 65280]) for i in range(0, len(raw), 65280)) + block(b"")
    assert gzip.decompress(blob) == raw
    p = tmp_path / "in.fastq.bgz"
    p.write_bytes(blob)
    for threads in (1, 4):
        text, (mode, _, _) = _text_and_mode(p, threads)
        assert mode == 1 and text == raw
    p.write_bytes(blob[:len(blob) // 2])
    with pytest.raises(ValueError, match="input"):
        _text_and_mode(p, 4)


def test_chunk_parallel_inflate_random_damage(tmp_path, big_text, monkeypatch):
    """Random damage (flipped bits, cuts, zeroed and inserted stretches) to one-member and many-member files:
    wherever python's gzip gives an error the reader gives one too, otherwise the same text; never a crash or a hang."""
    import random
    raw = big_text[:big_text.index(b"\n@r", 2500000) + 1]
    one = gzip.compress(raw, 6)
    cuts = [0] + [raw.index(b"\n@r", c) + 1 for c in range(250000, len(raw) - 250000, 250000)] + [len(raw)]
    many = b"".join(gzip.compress(raw[a:b], 6) for a, b in zip(cuts[:-1], cuts[1:]))
    monkeypatch.setenv("ORC_PGZ_MIN", "100000")
    rnd = random.Random(20)
    p = tmp_path / "d.fastq.gz"
    refused = 0
    for trial in range(40):
        data = bytearray(rnd.choice([one, many]))
        mode = rnd.choice(["flip", "cut", "zero", "insert"])
        if mode == "flip":
            data[rnd.randrange(len(data))] ^= 1 << rnd.randrange(8)
        elif mode == "cut":
            del data[rnd.randrange(100, len(data)):]
        elif mode == "zero":
            a = rnd.randrange(len(data) - 1)
            b = min(len(data), a + rnd.randint(1, 3000))
            data[a:b] = bytes(b - a)
        else:
            a = rnd.randrange(len(data))
            data[a:a] = bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 50)))
        monkeypatch.setenv("ORC_PGZ_CHUNK", str(rnd.choice([30000, 1 << 17, 1 << 20])))
        p.write_bytes(bytes(data))
        try:
            want = gzip.decompress(bytes(data))
        except Exception:
            want = None
        if want is None:
            refused += 1
            with pytest.raises(ValueError):
                _text_and_mode(p, rnd.choice([2, 5]))
        else:
            assert _text_and_mode(p, rnd.choice([2, 5]))[0] == want, (trial, mode)
    assert refused >= 30


def test_reader_closed_in_the_middle_of_a_chunk_parallel_stream(tmp_path, big_text, monkeypatch):
    """Closing the reader while the pool is still inflating (an error further down the pipeline) returns at once."""
    import time
    monkeypatch.setenv("ORC_PGZ_MIN", "100000")
    monkeypatch.setenv("ORC_PGZ_CHUNK", str(1 << 17))
    p = tmp_path / "in.fastq.gz"
    p.write_bytes(gzip.compress(big_text, 1))
    for take in (0, 2):
        rd = F.FastqReader(str(p), max_reads=2000, max_bytes=1 << 21, keep=2, ahead=2, pinned=False, threads=4)
        it = iter(rd)
        for _ in range(take):
            assert next(it).n_reads > 0
        assert rd.inflate_mode()[0] == 2
        t0 = time.time()
        rd.close()
        assert time.time() - t0 < 5.0
