"""Seeded synthetic ONT-like reads for the two-round demultiplex.

The reference ships no fixtures, so every test and benchmark input comes from here.
Template = SP5[x] + insert + SP27rc[y]; iid sequencing errors; random truncation, missing
adapters, reverse-complemented reads and the odd N.  Deterministic for a given
(seed, n_reads, len_min, len_max): reads are generated in fixed chunks of CHUNK reads,
chunk c with PCG64(SeedSequence([seed, c])).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import m13

CHUNK = 32768
_ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)

P_SUB, P_INS, P_DEL = 0.015, 0.010, 0.015
P_TRUNC, P_NOADAPTER, P_RC, P_N = 0.15, 0.05, 0.10, 0.005


@dataclass
class ReadSet:
    """A batch of reads in the layout the C ABI takes (include/orcdemux.h orc_batch)."""
    seq: np.ndarray           # uint8 ASCII, reads back to back
    qual: np.ndarray          # uint8 ASCII
    offsets: np.ndarray       # uint64 [n]
    lengths: np.ndarray       # uint32 [n]
    names: np.ndarray         # uint8 blob of header lines (no '@', no newline)
    name_offsets: np.ndarray  # uint64 [n+1]
    truth: dict               # sp5, sp27 (1..12, 0 = absent), rc, trunc5, trunc3

    @property
    def n_reads(self) -> int:
        return int(self.lengths.shape[0])

    def read(self, r: int):
        o, n = int(self.offsets[r]), int(self.lengths[r])
        a, b = int(self.name_offsets[r]), int(self.name_offsets[r + 1])
        return (self.names[a:b].tobytes().decode(), self.seq[o:o + n].tobytes().decode(),
                self.qual[o:o + n].tobytes().decode())

    def to_fastq_bytes(self) -> bytes:
        out = []
        for r in range(self.n_reads):
            nm, s, q = self.read(r)
            out.append("@%s\n%s\n+\n%s\n" % (nm, s, q))
        return "".join(out).encode()


def _codes(seqs):
    lut = np.zeros(256, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    return np.stack([lut[np.frombuffer(s.encode(), dtype=np.uint8)] for s in seqs])


def _chunk(seed: int, chunk: int, first: int, n: int, len_min: int, len_max: int,
           front, back, anchored_index=None):
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, chunk])))
    n_front, n_back = front.shape[0], back.shape[0]
    L5f, L27f = front.shape[1], back.shape[1]
    x = rng.integers(1, n_front + 1, n)
    y = rng.integers(1, n_back + 1, n)
    total = rng.integers(len_min, len_max + 1, n)
    no5 = rng.random(n) < P_NOADAPTER
    no27 = rng.random(n) < P_NOADAPTER
    t5 = np.where(rng.random(n) < P_TRUNC, rng.integers(1, 41, n), 0)
    t3 = np.where(rng.random(n) < P_TRUNC, rng.integers(1, 41, n), 0)
    is_rc = rng.random(n) < P_RC
    has_n = rng.random(n) < P_N
    if anchored_index is not None:
        # BASELINE config 4: the bare 17-nt index sits at read offset 0, no flanks, no truncation
        t5[:] = 0
        is_rc[:] = False
        no5[:] = False
    L5 = np.where(no5, 0, L5f)
    L27 = np.where(no27, 0, L27f)
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(total, out=starts[1:])
    T = int(starts[-1])
    tpl = rng.integers(0, 4, T, dtype=np.uint8)
    ar5 = np.arange(L5f, dtype=np.int64)
    ar27 = np.arange(L27f, dtype=np.int64)
    h5 = np.flatnonzero(~no5)
    tpl[(starts[h5][:, None] + ar5[None, :]).ravel()] = front[x[h5] - 1].ravel()
    h27 = np.flatnonzero(~no27)
    tpl[((starts[h27 + 1] - L27f)[:, None] + ar27[None, :]).ravel()] = back[y[h27] - 1].ravel()
    # iid errors over the whole template
    u = rng.random(T, dtype=np.float32)
    sub = u < P_SUB
    ins = (u >= P_SUB) & (u < P_SUB + P_INS)
    dele = (u >= P_SUB + P_INS) & (u < P_SUB + P_INS + P_DEL)
    nsub = int(sub.sum())
    tpl[sub] = (tpl[sub] + rng.integers(1, 4, nsub, dtype=np.uint8)) & 3
    counts = np.ones(T, dtype=np.int8)
    counts[ins] = 2
    counts[dele] = 0
    out = np.repeat(tpl, counts)
    csum = np.cumsum(counts, dtype=np.int64)
    ins_pos = csum[ins] - 1                       # the second copy is the inserted base
    out[ins_pos] = rng.integers(0, 4, ins_pos.shape[0], dtype=np.uint8)
    new_starts = np.zeros(n + 1, dtype=np.int64)
    new_starts[1:] = csum[starts[1:] - 1]
    lens = np.diff(new_starts)
    # truncation
    t5 = np.minimum(t5, lens // 2)
    t3 = np.minimum(t3, lens // 2)
    rid = np.repeat(np.arange(n, dtype=np.int64), lens)
    pos = np.arange(out.shape[0], dtype=np.int64) - new_starts[rid]
    keep = (pos >= t5[rid]) & (pos < (lens - t3)[rid])
    out = out[keep]
    lens = lens - t5 - t3
    fs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=fs[1:])
    # reverse complement
    rid = np.repeat(np.arange(n, dtype=np.int64), lens)
    pos = np.arange(out.shape[0], dtype=np.int64) - fs[rid]
    rcm = is_rc[rid]
    src = np.where(rcm, fs[rid] + lens[rid] - 1 - pos, fs[rid] + pos)
    out = np.where(rcm, 3 - out[src], out[src]).astype(np.uint8)
    # one N
    hn = np.flatnonzero(has_n & (lens > 0))
    out[fs[hn] + rng.integers(0, 1 << 30, hn.shape[0]) % lens[hn]] = 4
    seq = _ASCII[out]
    qual = (rng.integers(5, 41, out.shape[0], dtype=np.uint8) + 33).astype(np.uint8)
    names = []
    for i in range(n):
        g = first + i
        names.append("r%d ch=%d" % (g, g % 512) if g % 7 == 0 else "r%d" % g)
    truth = dict(sp5=np.where(no5, 0, x).astype(np.int32), sp27=np.where(no27, 0, y).astype(np.int32),
                 rc=is_rc.astype(np.uint8), trunc5=t5.astype(np.int32), trunc3=t3.astype(np.int32))
    return seq, qual, lens.astype(np.uint32), names, truth


def _chunk_star(args):
    return _chunk(*args)


def generate(n_reads: int, len_min: int = 300, len_max: int = 900, seed: int = 1002,
             anchored: bool = False, workers: int = 1) -> ReadSet:
    """SURVEY 8(d) read model.  anchored=True builds BASELINE config 4 inputs (bare index at offset 0).
    workers > 1 generates the chunks in a process pool (same bytes as workers == 1)."""
    if anchored:
        front = _codes([s for _, s in m13.variable_all()])
        back = np.zeros((1, 0), dtype=np.uint8)                  # no 3' adapter
    else:
        front = _codes([s for _, s in m13.sp5_forward()])
        back = _codes([s for _, s in m13.sp27_reverse_rc()])
    seqs, quals, lens, names, truths = [], [], [], [], []
    jobs = []
    first = 0
    c = 0
    while first < n_reads:
        n = min(CHUNK, n_reads - first)
        jobs.append((seed, c, first, n, len_min, len_max, front, back, True if anchored else None))
        first += n
        c += 1
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            parts = pool.map(_chunk_star, jobs)
    else:
        parts = [_chunk(*j) for j in jobs]
    for s, q, l, nm, t in parts:
        seqs.append(s); quals.append(q); lens.append(l); names.extend(nm); truths.append(t)
    seq = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    qual = np.concatenate(quals) if quals else np.zeros(0, np.uint8)
    lengths = np.concatenate(lens) if lens else np.zeros(0, np.uint32)
    offsets = np.zeros(n_reads, dtype=np.uint64)
    if n_reads > 1:
        offsets[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
    nb = [x.encode() for x in names]
    name_offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    if n_reads:
        name_offsets[1:] = np.cumsum([len(b) for b in nb], dtype=np.uint64)
    name_blob = np.frombuffer(b"".join(nb), dtype=np.uint8).copy() if nb else np.zeros(0, np.uint8)
    truth = {k: np.concatenate([t[k] for t in truths]) for k in truths[0]} if truths else {}
    return ReadSet(seq, qual, offsets, lengths, name_blob, name_offsets, truth)


def from_records(records) -> ReadSet:
    """Build a ReadSet from [(name, seq, qual)] (hand-written test cases)."""
    seq = np.frombuffer("".join(r[1] for r in records).encode(), dtype=np.uint8).copy()
    qual = np.frombuffer("".join(r[2] for r in records).encode(), dtype=np.uint8).copy()
    lengths = np.array([len(r[1]) for r in records], dtype=np.uint32)
    offsets = np.zeros(len(records), dtype=np.uint64)
    if len(records) > 1:
        offsets[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
    nb = [r[0].encode() for r in records]
    name_offsets = np.zeros(len(records) + 1, dtype=np.uint64)
    if records:
        name_offsets[1:] = np.cumsum([len(b) for b in nb], dtype=np.uint64)
    names = np.frombuffer(b"".join(nb), dtype=np.uint8).copy() if nb else np.zeros(0, np.uint8)
    return ReadSet(seq, qual, offsets, lengths, names, name_offsets, {})
