"""The block decoder of the chunk-parallel inflate (csrc/orc_pgz.h, through orc_gunzip_file) on DEFLATE streams
written bit by bit here -- the corners of the format that zlib's own compressor never or rarely produces -- and on
what zlib's compressor produces under every strategy.  The checker is zlib's inflate: the same bytes where it
accepts (inflate.c / inftrees.c), a refusal where it refuses.  Runs on the CPU."""
import os
import random
import zlib

import pytest

from orcdemux import fastq as F


class _Bits:
    def __init__(self):
        self.bits = []

    def put(self, v, n):                    # plain fields: least significant bit first
        self.bits.extend((v >> i) & 1 for i in range(n))

    def code(self, c, n):                   # Huffman codes: most significant bit first
        self.bits.extend((c >> i) & 1 for i in range(n - 1, -1, -1))

    def bytes(self):
        out = bytearray()
        for i in range(0, len(self.bits), 8):
            out.append(sum(b << j for j, b in enumerate(self.bits[i:i + 8])))
        return bytes(out)


def _canonical(lens):
    count = [0] * 16
    for l in lens:
        count[l] += 1
    count[0] = 0
    nxt, c = [0] * 16, 0
    for l in range(1, 16):
        c = (c + count[l - 1]) << 1
        nxt[l] = c
    codes = [0] * len(lens)
    for i, l in enumerate(lens):
        if l:
            codes[i] = nxt[l]
            nxt[l] += 1
    return codes


def _dynamic_block(bw, final, litlens, distlens, syms):
    """syms: ('L', byte) | ('M', length symbol, extra value, extra bits, distance symbol, extra value, extra bits) | ('E',)"""
    bw.put(final, 1)
    bw.put(2, 2)
    bw.put(len(litlens) - 257, 5)
    bw.put(len(distlens) - 1, 5)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    cl = [4] * 16 + [0, 0, 0]               # every length 0..15 a 4-bit code (complete), no repeat codes
    bw.put(19 - 4, 4)
    for o in order:
        bw.put(cl[o], 3)
    clcodes = _canonical(cl)
    for l in list(litlens) + list(distlens):
        bw.code(clcodes[l], 4)
    lc, dc = _canonical(litlens), _canonical(distlens)
    for s in syms:
        if s[0] == "L":
            bw.code(lc[s[1]], litlens[s[1]])
        elif s[0] == "E":
            bw.code(lc[256], litlens[256])
        else:
            _, ls, ev, eb, ds, dv, db = s
            bw.code(lc[ls], litlens[ls])
            bw.put(ev, eb)
            bw.code(dc[ds], distlens[ds])
            bw.put(dv, db)


def _gz(raw, data):
    return (bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3]) + raw + zlib.crc32(data).to_bytes(4, "little") +
            (len(data) & 0xffffffff).to_bytes(4, "little"))


def _check(tmp_path, raw, chunks=(4096, 1 << 20), threads=(1, 4)):
    data = zlib.decompress(raw, -15)
    p = tmp_path / "c.gz"
    p.write_bytes(_gz(raw, data))
    for chunk in chunks:
        for t in threads:
            assert F.gunzip_file(str(p), len(data) + 16, t, chunk) == data, (chunk, t)
    return data


def test_crafted_streams(tmp_path):
    rnd = random.Random(1)
    # one distance code of one bit (an incomplete set inftrees.c lets pass), used by every match
    lit = [0] * 258
    lit[ord("a")] = lit[ord("b")] = lit[256] = lit[257] = 2
    bw = _Bits()
    syms = [("L", ord("a")), ("L", ord("b"))] + [("M", 257, 0, 0, 0, 0, 0)] * 500 + [("L", ord("a"))] * 3 + [("E",)]
    for i in range(40):
        _dynamic_block(bw, 1 if i == 39 else 0, lit, [1], syms)
    assert len(_check(tmp_path, bw.bytes())) == 40 * 1505
    # no distance code at all (one code of zero bits), 9-bit literals, a 1-bit end-of-block code
    lit = [9] * 256 + [1]
    bw = _Bits()
    for i in range(30):
        _dynamic_block(bw, 1 if i == 29 else 0, lit, [0], [("L", rnd.randrange(256)) for _ in range(3000)] + [("E",)])
    _check(tmp_path, bw.bytes())
    # codes of 1, 2, ... 14, 15, 15 bits: second-level tables of every depth
    lit = [0] * 257
    for i, l in enumerate(list(range(1, 15)) + [15]):
        lit[65 + i] = l
    lit[256] = 15
    bw = _Bits()
    for i in range(30):
        _dynamic_block(bw, 1 if i == 29 else 0, lit, [0],
                       [("L", 65 + min(14, int(rnd.expovariate(0.7)))) for _ in range(5000)] + [("E",)])
    _check(tmp_path, bw.bytes())
    # length 258 from 27 577 bytes back (13 extra bits) and from 1 byte back
    lit = [0] * 286
    lit[ord("x")], lit[256], lit[285] = 1, 2, 2
    dist = [0] * 30
    dist[29] = dist[0] = 1
    bw = _Bits()
    syms = [("L", ord("x"))] * 30000 + [("M", 285, 0, 0, 29, 3000, 13), ("M", 285, 0, 0, 0, 0, 0)] * 20 + [("E",)]
    for i in range(12):
        _dynamic_block(bw, 1 if i == 11 else 0, lit, dist, syms)
    _check(tmp_path, bw.bytes())


def test_crafted_streams_zlib_refuses(tmp_path):
    def refused(raw, what):
        with pytest.raises(zlib.error, match=what):
            zlib.decompress(raw, -15)
        p = tmp_path / "bad.gz"
        p.write_bytes(_gz(raw, b"") + bytes(64))
        with pytest.raises(ValueError, match=what):
            F.gunzip_file(str(p), 1 << 20, 2, 4096)

    lit = [0] * 257
    lit[65] = lit[66] = lit[256] = 1
    bw = _Bits()
    _dynamic_block(bw, 1, lit, [0], [("L", 65), ("E",)])
    refused(bw.bytes(), "invalid literal/lengths set")             # over-subscribed
    lit = [0] * 257
    lit[65] = lit[66] = 1
    bw = _Bits()
    _dynamic_block(bw, 1, lit, [0], [("L", 65)])
    refused(bw.bytes(), "missing end-of-block")
    lit = [0] * 257
    lit[65] = lit[66] = lit[256] = 2
    bw = _Bits()
    _dynamic_block(bw, 1, lit, [0], [("L", 65), ("E",)])
    refused(bw.bytes(), "invalid literal/lengths set")             # incomplete
    lit = [0] * 258
    lit[65] = lit[66] = lit[256] = lit[257] = 2
    dist = [0] * 30
    dist[10] = dist[0] = 1
    bw = _Bits()
    _dynamic_block(bw, 1, lit, dist, [("L", 65), ("M", 257, 0, 0, 10, 0, 4), ("E",)])
    refused(bw.bytes(), "invalid distance too far back")
    bw = _Bits()
    bw.put(1, 1)
    bw.put(3, 2)
    refused(bw.bytes() + bytes(8), "invalid block type")
    refused(bytes([1, 5, 0, 0xfb, 0xff]) + b"hello", "invalid stored block lengths")


def test_what_zlib_writes_under_every_strategy(tmp_path):
    """Levels 0-9, memLevel 1-9, the five strategies, flushes thrown in, several members; text, random bytes,
    runs, zeros; chunks from 4 KB (most blocks longer than a chunk) to 1 MiB."""
    rnd = random.Random(3)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(200)),
                                              bytes(33 + min(40, int(rnd.expovariate(0.1))) for _ in range(200)))
                    for i in range(6000))

    def some(kind, n):
        if kind == 0:
            o = rnd.randrange(max(1, len(text) - n))
            return text[o:o + n]
        if kind == 1:
            return os.urandom(n)
        if kind == 2:
            return bytes(rnd.choice(b"ACGT") for _ in range(997)) * (n // 997)
        return bytes(n)

    p = tmp_path / "s.gz"
    for trial in range(40):
        blob, want = b"", b""
        for _ in range(rnd.choice([1, 1, 2, 7])):
            d = some(rnd.randrange(4), rnd.choice([0, 1, 3000, 200000, 1500000]))
            co = zlib.compressobj(rnd.randint(0, 9), zlib.DEFLATED, 31, rnd.randint(1, 9), rnd.randrange(5))
            step = rnd.choice([len(d) or 1, 65536])
            for i in range(0, len(d), step):
                blob += co.compress(d[i:i + step])
                if rnd.random() < 0.2:
                    blob += co.flush(rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))
            blob += co.flush()
            want += d
        if len(blob) < 64:
            continue
        p.write_bytes(blob)
        got = F.gunzip_file(str(p), len(want) + 16, rnd.choice([1, 2, 5]), rnd.choice([4096, 20000, 1 << 16, 1 << 20]))
        assert got == want, trial
