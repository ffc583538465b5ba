"""Pairwise edit distance (what amplicon_sorter asks of edlib,
amplicon_sorter.py:225-235 NW, :838-849 HW).  The oracle is the textbook recurrence
(oracle/edit_oracle.c) -- pinned by definition; checked here against hand-known answers and against
the 64-row block formulation the kernel uses (csrc/orc_edit.cuh, run on the host)."""
import ctypes as C
import random

import numpy as np
import pytest

import helpers as H
import oracle
from orcdemux import distance as D


def _block(q: bytes, t: bytes, mode: int) -> int:
    L = H.hostsim()
    L.hostsim_edit_distance.restype = C.c_uint32
    L.hostsim_edit_distance.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int]
    return int(L.hostsim_edit_distance(q, len(q), t, len(t), mode))


def _cases(rnd, n):
    out = []
    for _ in range(n):
        m = rnd.choice([0, 1, 5, 63, 64, 65, 127, 128, 129, 200, 300])
        a = bytes(rnd.choice(b"\x00\x01\x02\x03\x04") for _ in range(m))
        b = bytearray(a)
        for _ in range(rnd.randint(0, 12)):                      # edits
            if b and rnd.random() < 0.6:
                p = rnd.randrange(len(b))
                r = rnd.random()
                if r < 0.4:
                    b[p] = rnd.choice(b"\x00\x01\x02\x03")
                elif r < 0.7:
                    del b[p]
                else:
                    b.insert(p, rnd.choice(b"\x00\x01\x02\x03"))
        pad = bytes(rnd.choice(b"\x00\x01\x02\x03") for _ in range(rnd.choice([0, 0, 10, 70])))
        out.append((a, pad + bytes(b) + pad[::-1]))
    return out


def test_oracle_known_answers():
    blob, off, ln = D.pack([b"kitten", b"sitting", b"ACGT", b"TTACGTTT", b"", b"A", b"ACGT"])
    assert list(oracle.edit_distances(blob, off, ln, [0, 2, 4, 2, 2], [1, 3, 5, 6, 4], "NW")) == [3, 4, 1, 0, 4]
    assert list(oracle.edit_distances(blob, off, ln, [0, 2, 4, 1], [1, 3, 5, 0], "HW")) == [2, 0, 0, 2]
    assert list(D.identities(np.array([3, 4]), ln, [0, 2], [1, 3])) == [round(1 - 3 / 7, 3), 0.5]


def test_block_formulation_equals_the_recurrence():
    rnd = random.Random(7)
    for q, t in _cases(rnd, 300):
        if len(q) > len(t):
            q, t = t, q
        blob, off, ln = D.pack([q, t])
        for mode, name in ((0, "NW"), (1, "HW")):
            exp = int(oracle.edit_distances(blob, off, ln, [0], [1], name, n_threads=1)[0])
            assert _block(q, t, mode) == exp, (name, len(q), len(t))


@pytest.mark.gpu
def test_kernel_equals_the_recurrence():
    rnd = random.Random(11)
    seqs = []
    for q, t in _cases(rnd, 200):
        seqs += [q, t]
    mk = lambda n: bytes(rnd.choice(b"ACGT") for _ in range(n))
    base = mk(5000)
    # lengths around the lane / block-count boundaries: 1, 2 and 4 blocks per lane
    # every lanes-per-pair class (5, 6, 7, 8, 10, 16, 32 lanes), then 2 and 4 blocks per lane
    for n in (320, 321, 384, 385, 448, 449, 512, 513, 600, 640, 641, 1000, 1024, 1025, 2047, 2048, 2049, 3000, 4096, 4097, 5000):
        s = bytearray(base[:n])
        for _ in range(40):
            s[rnd.randrange(n)] = rnd.choice(b"ACGTN")
        seqs += [base[:n], bytes(s)]
    seqs += [b"", b"", b"N" * 70, b"N" * 64 + b"ACGT"]
    blob, off, ln = D.pack(seqs)
    n = len(seqs)
    pa = np.arange(0, n, 2, dtype=np.uint32)
    pb = pa + 1
    extra_a = np.array([rnd.randrange(n) for _ in range(300)], dtype=np.uint32)
    extra_b = np.array([rnd.randrange(n) for _ in range(300)], dtype=np.uint32)
    pa, pb = np.concatenate([pa, extra_a]), np.concatenate([pb, extra_b])
    # five ACGTN symbols plus the three of the synthetic cases' alphabet would exceed 8: remap those
    blob = blob.copy()
    for i, c in enumerate(b"ACGTN"):
        blob[blob == i] = c
    for name in ("NW", "HW"):
        got = D.edit_distances(blob, off, ln, pa, pb, name)
        exp = oracle.edit_distances(blob, off, ln, pa, pb, name)
        bad = np.flatnonzero(got != exp)
        assert bad.size == 0, (name, bad[:5], got[bad[:5]], exp[bad[:5]], ln[pa[bad[:5]]], ln[pb[bad[:5]]])


@pytest.mark.gpu
def test_kernel_refuses_what_it_cannot_do():
    from orcdemux.engine import OrcError
    blob, off, ln = D.pack([bytes(range(65, 75)), b"ACGT"])
    with pytest.raises(OrcError, match="distinct characters"):
        D.edit_distances(blob, off, ln, [0], [1])
    blob, off, ln = D.pack([b"A" * 9000, b"C" * 9000])
    with pytest.raises(OrcError, match="longer than 8192"):
        D.edit_distances(blob, off, ln, [0], [1])
    blob, off, ln = D.pack([b"A" * 9000, b"ACGT"])                # only the query is limited
    assert list(D.edit_distances(blob, off, ln, [0], [1], "NW")) == [8999]
    with pytest.raises(OrcError, match="out of range"):
        D.edit_distances(blob, off, ln, [0], [5])


@pytest.mark.gpu
def test_similarity_worker_semantics():
    """amplicon_sorter.py:777-808: threshold on the forward identity, reverse-complement retry below 0.5."""
    from orcdemux import synth
    rnd = random.Random(3)
    fam = []
    for _ in range(4):                                   # four families of similar reads
        base = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(400, 600)))
        for _ in range(6):
            s = bytearray(base)
            for _ in range(rnd.randint(0, 40)):
                s[rnd.randrange(len(s))] = rnd.choice(b"ACGT")
            fam.append(bytes(s) if rnd.random() < 0.7 else D.compl_reverse(bytes(s)))
    pa, pb = D.all_pairs(len(fam))
    got = D.similarity(fam, pa, pb, similar_genes=85.0)
    exp = []
    for a, b in zip(pa.tolist(), pb.tolist()):
        def iden(x, y):
            blob, off, ln = D.pack([x, y])
            d = int(oracle.edit_distances(blob, off, ln, [0], [1], "NW", n_threads=1)[0])
            return round(1 - d / max(len(x), len(y)), 3)
        i = iden(fam[a], fam[b])
        if i >= 0.85:
            exp.append((a, b, i, False))
        elif i < 0.5:
            i = iden(fam[a], D.compl_reverse(fam[b]))
            if i >= 0.85:
                exp.append((a, b, i, True))
    assert got == exp and any(r[3] for r in got) and any(not r[3] for r in got)
    assert D.compl_reverse(b"AACGTN") == b"NACGTT"
