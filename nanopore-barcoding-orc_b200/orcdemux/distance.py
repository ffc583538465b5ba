"""Pairwise edit distances on the GPU: the call amplicon_sorter makes to edlib for every pair of
reads it compares (/root/reference/scripts/auxiliary_code/amplicon_sorter.py:225-235 `distance`,
:838-849 `distance_finetune`), behind the C ABI entry `orc_edit_distances` (include/orcdemux.h).

    iden = round(1 - edlib.align(shorter, longer, task='distance', mode=mode)['editDistance'] / len(longer), 3)

All arithmetic is in liborcdemux.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import lib as _lib
from .engine import OrcError

MODES = {"NW": 0, "HW": 1}


def pack(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(blob, offsets, lengths) of a list of sequences (bytes or str)."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    lengths = np.array([len(b) for b in bs], dtype=np.uint32)
    offsets = np.zeros(len(bs), dtype=np.uint64)
    if len(bs) > 1:
        offsets[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return blob, offsets, lengths


def edit_distances(blob, offsets, lengths, pair_a, pair_b, mode: str = "NW", device: int = 0, with_time: bool = False):
    """editDistance of every pair (pair_a[k], pair_b[k]); the shorter sequence is the query (edlib's
    first argument), on equal lengths pair_a[k] is."""
    L = _lib.load()
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    if pa.shape != pb.shape:
        raise ValueError("pair_a and pair_b differ in length")
    out = np.zeros(pa.shape[0], dtype=np.uint32)
    ms = C.c_float(0.0)
    err = C.create_string_buffer(256)
    rc = L.orc_edit_distances(device, blob.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, lengths.shape[0],
                              pa.ctypes.data, pb.ctypes.data, pa.shape[0], MODES[mode], out.ctypes.data,
                              C.byref(ms), err, 256)
    if rc != _lib.ORC_OK:
        raise OrcError("orc_edit_distances failed (%d): %s" % (rc, err.value.decode(errors="replace")))
    return (out, float(ms.value)) if with_time else out


def identities(dist: np.ndarray, lengths: np.ndarray, pair_a, pair_b) -> np.ndarray:
    """amplicon_sorter's similarity: round(1 - distance / len(longer), 3)."""
    la = lengths[np.asarray(pair_a)].astype(np.float64)
    lb = lengths[np.asarray(pair_b)].astype(np.float64)
    x = 1.0 - dist / np.maximum(np.maximum(la, lb), 1.0)
    r = np.round(x, 3)
    # Python's round() rounds the exact binary value, numpy rounds x * 1000: they can differ next to
    # a tie, so those few values go through Python's
    y = x * 1000.0
    for i in np.flatnonzero(np.abs(y - np.floor(y) - 0.5) < 1e-6):
        r[i] = round(float(x[i]), 3)
    return r


def all_pairs(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Every unordered pair i < j of n sequences (amplicon_sorter's compare-all mode)."""
    a, b = np.triu_indices(n, k=1)
    return a.astype(np.uint32), b.astype(np.uint32)


_COMPL = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ATCGRYKMSW", b"TAGCYRMKSW"):        # amplicon_sorter.py:237-242 compl_reverse
    _COMPL[_a] = _b


def compl_reverse(seq: bytes) -> bytes:
    """amplicon_sorter.py:237-242: reverse, then complement ATCGRYKMSW (other characters unchanged)."""
    return _COMPL[np.frombuffer(seq, dtype=np.uint8)[::-1]].tobytes()


def similarity(seqs: Sequence[bytes], pair_a, pair_b, similar_genes: float = 80.0, device: int = 0):
    """What amplicon_sorter's similarity() worker writes for the listed pairs
    (amplicon_sorter.py:777-808): (i, j, iden, reversed) for every pair whose identity reaches
    similar_genes per cent; a pair below 0.5 is tried again against the reverse complement of
    its second sequence and kept, flagged 'reverse', if that reaches the threshold."""
    n = len(seqs)
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    blob, off, ln = pack(bs + [compl_reverse(b) for b in bs])     # sequence n + i = reverse complement of i
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    thr = similar_genes / 100.0
    iden = identities(edit_distances(blob, off, ln, pa, pb, "NW", device), ln, pa, pb)
    out = [(int(a), int(b), float(i), False) for a, b, i in zip(pa[iden >= thr], pb[iden >= thr], iden[iden >= thr])]
    low = np.flatnonzero(iden < 0.5)
    if low.size:
        ra, rb = pa[low], pb[low] + np.uint32(n)
        riden = identities(edit_distances(blob, off, ln, ra, rb, "NW", device), ln, ra, rb)
        keep = riden >= thr
        out += [(int(a), int(b), float(i), True) for a, b, i in zip(pa[low][keep], pb[low][keep], riden[keep])]
    out.sort(key=lambda r: (r[0], r[1]))
    return out
