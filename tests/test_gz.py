"""The device-side gzip encoder (csrc/orc_gz.cuh), run on the CPU through tests/gzsim.cpp: every member must be
what zlib and Python's gzip module accept, decompress to the bin's bytes, and carry the library's "OC" size
field (orc_io.cpp finds members by it).  The `-m gpu` test in tests/test_gpu_parity.py repeats the round trip
on the kernels' output."""
import ctypes as C
import gzip
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

import helpers as H
from orcdemux import synth

_SO = os.path.join(H.ROOT, "tests", "_build", "libgzsim.so")
_SRC = os.path.join(H.ROOT, "tests", "gzsim.cpp")
_HDR = os.path.join(H.PKG, "csrc", "orc_gz.cuh")


def _lib():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(_SRC), os.path.getmtime(_HDR)):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", _SO, _SRC], check=True)
    L = C.CDLL(_SO)
    L.gzsim_compress.restype = C.c_longlong
    L.gzsim_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    return L


def check_members(gz: bytes, gz_offsets, pieces):
    """Every member: header with the size field, zlib inflates it to the piece, CRC and ISIZE hold."""
    for m, piece in enumerate(pieces):
        a, b = int(gz_offsets[m]), int(gz_offsets[m + 1])
        if len(piece) == 0:
            assert a == b
            continue
        mem = gz[a:b]
        assert mem[:4] == b"\x1f\x8b\x08\x04" and mem[10:16] == b"\x08\x00OC\x04\x00"
        assert int.from_bytes(mem[16:20], "little") == b - a
        d = zlib.decompressobj(31)
        out = d.decompress(mem)
        assert d.eof and d.unused_data == b"" and out == piece, "member %d" % m
        assert int.from_bytes(mem[-8:-4], "little") == zlib.crc32(piece)
        assert int.from_bytes(mem[-4:], "little") == len(piece) & 0xFFFFFFFF
    whole = b"".join(gz[int(gz_offsets[m]):int(gz_offsets[m + 1])] for m in range(len(pieces)))
    assert gzip.decompress(whole) == b"".join(pieces) if whole else True      # members concatenate to a gzip file


def _compress(pieces):
    L = _lib()
    text = np.frombuffer(b"".join(pieces) + b"\0" * 64, dtype=np.uint8).copy()
    offs = np.zeros(len(pieces) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(p) for p in pieces])
    cap = (int(offs[-1]) * 2 + 4096 * (len(pieces) + 1) + 3) & ~3
    out = np.zeros(cap, dtype=np.uint8)
    gzo = np.zeros(len(pieces) + 1, dtype=np.uint64)
    lens = np.zeros(257, dtype=np.uint8)
    n = L.gzsim_compress(text.ctypes.data, offs.ctypes.data, len(pieces), out.ctypes.data, cap, gzo.ctypes.data, lens.ctypes.data)
    assert n >= 0 and n == int(gzo[-1])
    return out[:n].tobytes(), gzo, lens


def test_fastq_bins_round_trip_through_zlib():
    rs = synth.generate(3000, 300, 900, seed=11)
    text = rs.to_fastq_bytes()
    rnd = random.Random(5)
    cuts = sorted(rnd.sample(range(1, len(text)), 40))
    cuts = [0] + cuts[:10] + [cuts[10]] * 3 + cuts[10:] + [len(text)]         # a few empty bins among them
    pieces = [text[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    gz, gzo, lens = _compress(pieces)
    check_members(gz, gzo, pieces)
    assert lens[ord("A")] <= 4 and lens[256] >= 1                            # bases get short codes
    ratio = len(gz) / len(text)
    zl = len(zlib.compress(text, 5)) / len(text)
    assert ratio < 0.62 and ratio < zl * 1.12, (ratio, zl)                   # literals-only costs little on FASTQ


def test_sizes_around_chunk_and_word_boundaries():
    rnd = random.Random(7)
    pieces = []
    for n in [1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 500, 511, 512, 513, 1023, 1024, 1025, 1536, 2047, 2048, 2049, 4096, 4097, 10000]:
        pieces.append(bytes(rnd.choice(b"ACGTN@+\n!#%I") for _ in range(n)))
        pieces.append(b"")
    gz, gzo, _ = _compress(pieces)
    check_members(gz, gzo, pieces)


def test_all_byte_values_and_long_codes():
    rnd = random.Random(9)
    # every byte value, with weights that fall by a factor of two each: a Huffman tree deeper than 15 before
    # the limit is applied
    data = bytearray()
    for b in range(256):
        data += bytes([b]) * max(1, (1 << 22) >> min(b, 22))
    rnd.shuffle(data)
    data = bytes(data)
    pieces = [data[:100000], data[100000:100007], data[100007:]]
    gz, gzo, lens = _compress(pieces)
    assert int(lens.max()) <= 15 and int((lens[:256] > 0).sum()) == 256
    check_members(gz, gzo, pieces)
    # a batch with a single distinct byte, and an empty batch
    gz, gzo, lens = _compress([b"A" * 5000, b"", b"A"])
    check_members(gz, gzo, [b"A" * 5000, b"", b"A"])
    gz, gzo, lens = _compress([b"", b""])
    assert len(gz) == 0


def test_one_large_member():
    rnd = np.random.default_rng(3)
    data = rnd.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=3_000_000).tobytes()
    gz, gzo, _ = _compress([data])
    check_members(gz, gzo, [data])
    assert len(gz) < 0.29 * len(data)                                         # two bits per base (one of the four takes three: the end-of-block code needs a leaf)


def test_flat_code_fits_nine_eighths(monkeypatch):
    """The fallback of orc_wait() when a batch's members outgrow their arena (gz_table_kernel flat = 1): 8- and
    9-bit codes whatever the bytes are, so a member never exceeds 9/8 of its text plus frame and block header."""
    monkeypatch.setenv("GZSIM_FLAT", "1")
    rnd = np.random.default_rng(5)
    data = rnd.integers(0, 256, size=400_000, dtype=np.uint8).tobytes()
    pieces = [data[:1000], data[1000:250_000], b"", data[250_000:]]
    gz, gzo, lens = _compress(pieces)
    assert int(lens.min()) >= 8 and int(lens.max()) <= 9
    check_members(gz, gzo, pieces)
    for m, piece in enumerate(pieces):
        if piece:
            assert int(gzo[m + 1]) - int(gzo[m]) <= len(piece) * 9 // 8 + 256


def test_writer_takes_members_as_they_are(tmp_path):
    """orc_writer_write_members (csrc/orc_io.cpp): the members of successive batches appended to .gz bin files as
    they are and inflated into plain files, empty bins left as valid empty .gz, the byte counts per bin taken from
    the members' ISIZE, a malformed member refused -- no GPU needed: the members come from the host build of the
    encoder."""
    from orcdemux import fastq as F
    rs = synth.generate(1500, 300, 900, seed=21)
    text = rs.to_fastq_bytes()
    whole = synth.generate(300, 300, 900, seed=22).to_fastq_bytes()      # bin 0 holds whole records: read back below
    cuts = [1000, 1000, 200000, len(text) // 2, len(text)]
    batch1 = [whole] + [text[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    batch2 = [b"", text[:5000], b"", b"", text[5000:9000]]
    paths = [str(tmp_path / "b0.fastq.gz"), str(tmp_path / "b1.fastq.gz"), str(tmp_path / "b2.fastq.gz"),
             str(tmp_path / "b3.fastq"), None]
    w = F.BinWriters(paths, 1, threads=3)

    class R:
        pass
    for pieces in (batch1, batch2):
        gz, gzo, _ = _compress(pieces)
        r = R()
        r.fastq = np.frombuffer(gz, dtype=np.uint8)
        r.bin_offsets = gzo
        w.wait(w.write_batch(r, members=True))
    bad = R()
    bad.fastq = np.frombuffer(b"not a gzip member at all, just thirty bytes..", dtype=np.uint8)
    bad.bin_offsets = np.array([0, 40, 40, 40, 40, 40], dtype=np.uint64)
    with pytest.raises(OSError):
        w.write_batch(bad, members=True)
    w.close()
    assert gzip.open(paths[0], "rb").read() == batch1[0] + batch2[0]
    assert gzip.open(paths[1], "rb").read() == batch1[1] + batch2[1]
    assert gzip.open(paths[2], "rb").read() == batch1[2] + batch2[2]
    assert open(paths[3], "rb").read() == batch1[3] + batch2[3]
    assert w.bytes_written[:4] == [len(batch1[i]) + len(batch2[i]) for i in range(4)]
    # what the writer left can be read back by the library's own reader (member-parallel inflate)
    rd = F.FastqReader(paths[0], max_reads=1 << 16, max_bytes=1 << 26, keep=1, ahead=1, threads=4)
    n = sum(tb.n_reads for tb in rd)
    rd.close()
    assert n == 300
