"""Shared helpers of the test-suite: build/load the host simulation of the device
arithmetic (tests/hostsim.cpp), run the oracle on a ReadSet, compare records."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nanopore-barcoding-orc_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from orcdemux import m13, synth  # noqa: E402

MATCH_DTYPE = oracle.MATCH_DTYPE
FIELDS = ["adapter", "is_rc", "ref_start", "ref_stop", "query_start", "query_stop", "score", "errors"]

_hostsim = None


def hostsim():
    global _hostsim
    if _hostsim is None:
        out = os.path.join(ROOT, "tests", "_build", "libhostsim.so")
        src = os.path.join(ROOT, "tests", "hostsim.cpp")
        deps = [src] + [os.path.join(PKG, "csrc", f) for f in ("orc_core.cuh", "orc_table.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                            "-I", os.path.join(PKG, "csrc"), "-o", out, src], check=True)
        _hostsim = C.CDLL(out)
        _hostsim.hostsim_demux.restype = C.c_int
    return _hostsim


def _cstrs(seqs):
    arr = (C.c_char_p * max(len(seqs), 1))()
    for i, s in enumerate(seqs):
        arr[i] = s.encode()
    return arr


def run_hostsim(rounds, rs, filter_mode=1, want_columns=False, indels=1):
    """rounds: [(sequences, type, e, O, rc)].  Returns (m0, m1, lo, len, rc, n_tasks)."""
    n = rs.n_reads
    m0 = np.zeros(n, dtype=MATCH_DTYPE)
    m1 = np.zeros(n, dtype=MATCH_DTYPE)
    lo = np.zeros(n, dtype=np.uint64)
    ln = np.zeros(n, dtype=np.uint32)
    rc = np.zeros(n, dtype=np.uint32)
    nt = np.zeros(2, dtype=np.uint64)
    ncol = np.zeros(2, dtype=np.uint64)
    err = C.create_string_buffer(256)
    r0 = rounds[0]
    r1 = rounds[1] if len(rounds) > 1 else rounds[0]
    a0, a1 = _cstrs(r0[0]), _cstrs(r1[0])
    seq = np.ascontiguousarray(rs.seq)
    ret = hostsim().hostsim_demux(
        C.c_int(len(rounds)),
        C.c_int(len(r0[0])), C.c_int(r0[1]), a0, C.c_double(r0[2]), C.c_int(r0[3]), C.c_int(r0[4]),
        C.c_int(len(r1[0])), C.c_int(r1[1]), a1, C.c_double(r1[2]), C.c_int(r1[3]), C.c_int(r1[4]),
        C.c_void_p(seq.ctypes.data), C.c_void_p(rs.offsets.ctypes.data), C.c_void_p(rs.lengths.ctypes.data),
        C.c_uint32(n), C.c_uint64(seq.shape[0]),
        C.c_void_p(m0.ctypes.data), C.c_void_p(m1.ctypes.data), C.c_void_p(lo.ctypes.data),
        C.c_void_p(ln.ctypes.data), C.c_void_p(rc.ctypes.data), C.c_void_p(nt.ctypes.data), err, C.c_int(256),
        C.c_int(filter_mode), C.c_void_p(ncol.ctypes.data), C.c_int(indels))
    if ret != 0:
        raise RuntimeError(err.value.decode())
    if want_columns:
        return m0, m1, lo, ln, rc, nt, ncol
    return m0, m1, lo, ln, rc, nt


def run_oracle(rounds, rs, n_threads=8, indels=True):
    sets = [(oracle.AdapterSet(r[0], r[1], r[2], r[3], indels=indels), r[4]) for r in rounds]
    return oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=n_threads)


def m13_rounds(e=0.1, overlap=3, rc=1):
    return [([s for _, s in m13.sp5_forward()], oracle.FRONT, e, overlap, rc),
            ([s for _, s in m13.sp27_reverse_rc()], oracle.BACK, e, overlap, rc)]


def diff_matches(a, b, limit=5):
    """Indices where two match-record arrays differ (all fields)."""
    bad = np.zeros(a.shape[0], dtype=bool)
    for f in FIELDS:
        bad |= a[f] != b[f]
    return np.flatnonzero(bad)[:limit], int(bad.sum())


def view_bytes(rs, lo, ln, rc):
    """Materialise the trimmed reads described by (lo, len, rc) views: list of (seq, qual) bytes."""
    comp = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"ACGTUMRWSYKVHDBN", b"TGCAAKYWSRMBDHVN"):
        comp[a] = b
        comp[a | 0x20] = b | 0x20
    out = []
    for i in range(lo.shape[0]):
        s = rs.seq[int(lo[i]):int(lo[i]) + int(ln[i])]
        q = rs.qual[int(lo[i]):int(lo[i]) + int(ln[i])]
        if rc[i] & 1:
            s = comp[s[::-1]]
            q = q[::-1]
        elif (int(rc[i]) >> 8) >= 2:
            # flipped in both rounds: dnaio's table sends U to A and A to T
            s = s.copy()
            s[s == ord("U")] = ord("T")
            s[s == ord("u")] = ord("t")
        out.append((s.tobytes(), q.tobytes()))
    return out
