"""ctypes binding of the C oracle (oracle/cutadapt_oracle.c) -- TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED *** -- see the header of cutadapt_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package; the product (nanopore-barcoding-orc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

FRONT, BACK, PREFIX, SUFFIX = 0, 1, 2, 3
ORA_MAX_ADAPTER = 256

MATCH_DTYPE = np.dtype([
    ("adapter", "<i4"), ("is_rc", "<i4"), ("ref_start", "<i4"), ("ref_stop", "<i4"),
    ("query_start", "<i4"), ("query_stop", "<i4"), ("score", "<i4"), ("errors", "<i4"),
])


class OraAdapter(C.Structure):
    _fields_ = [
        ("seq", C.c_char * (ORA_MAX_ADAPTER + 1)),
        ("m", C.c_int),
        ("type", C.c_int),
        ("max_error_rate", C.c_double),
        ("min_overlap", C.c_int),
        ("indels", C.c_int),
        ("adapter_wildcards", C.c_int),
        ("read_wildcards", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("cutadapt_oracle.c", "edit_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        assert L.oracle_sizeof_adapter() == C.sizeof(OraAdapter), "oracle struct layout drift"
        loc_args = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                    C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_locate.argtypes = loc_args
        L.oracle_locate.restype = C.c_int
        L.oracle_locate_unpruned.argtypes = loc_args
        L.oracle_locate_unpruned.restype = C.c_int
        L.oracle_affix_compare.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_double, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_affix_compare.restype = C.c_int
        L.oracle_adapter_init.argtypes = [C.POINTER(OraAdapter), C.c_char_p, C.c_int, C.c_double, C.c_int,
                                          C.c_int, C.c_int, C.c_int]
        L.oracle_adapter_init.restype = None
        L.oracle_adapter_match.argtypes = [C.POINTER(OraAdapter), C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_adapter_match.restype = C.c_int
        L.oracle_best_of.argtypes = [C.POINTER(OraAdapter), C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_best_of.restype = C.c_int
        L.oracle_demux_batch.argtypes = [
            C.c_int,
            C.POINTER(OraAdapter), C.c_int, C.c_int,
            C.POINTER(OraAdapter), C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_demux_batch.restype = C.c_int
        _lib = L
    return _lib


def locate(ref: str, query: str, max_error_rate: float, flags: int, min_overlap: int = 1,
           indel_cost: int = 1, wildcard_ref: bool = False, wildcard_query: bool = False,
           unpruned: bool = False):
    """Aligner(ref, max_error_rate, flags, ...).locate(query) -> 6-tuple or None."""
    out = (C.c_int * 6)()
    fn = lib().oracle_locate_unpruned if unpruned else lib().oracle_locate
    r = fn(ref.encode(), len(ref), query.encode(), len(query), max_error_rate, flags, min_overlap,
           indel_cost, int(wildcard_ref), int(wildcard_query), out)
    return tuple(out) if r == 1 else None


def affix_compare(ref, query, max_error_rate, min_overlap=1, wildcard_ref=False, wildcard_query=False,
                  suffix=False):
    out = (C.c_int * 6)()
    r = lib().oracle_affix_compare(ref.encode(), len(ref), query.encode(), len(query), max_error_rate,
                                   min_overlap, int(wildcard_ref), int(wildcard_query), int(suffix), out)
    return tuple(out) if r == 1 else None


class AdapterSet:
    """A file-ordered list of adapters of one type (what `-g file:X` / `-a file:X` builds)."""

    def __init__(self, sequences, where, max_errors=0.1, min_overlap=3, indels=True,
                 adapter_wildcards=True, read_wildcards=False, names=None):
        self.n = len(sequences)
        self.where = where
        self.names = list(names) if names is not None else [str(i + 1) for i in range(self.n)]
        self.arr = (OraAdapter * max(self.n, 1))()
        for i, s in enumerate(sequences):
            lib().oracle_adapter_init(C.byref(self.arr[i]), s.encode(), where, float(max_errors),
                                      int(min_overlap), int(indels), int(adapter_wildcards),
                                      int(read_wildcards))

    def best_of(self, query_upper: str):
        """MultipleAdapters / IndexedPrefixAdapters.match_to -> (adapter index, 6-tuple) or None."""
        out = (C.c_int * 6)()
        a = lib().oracle_best_of(self.arr, self.n, query_upper.encode(), len(query_upper), out)
        return (a, tuple(out)) if a >= 0 else None

    def match(self, idx: int, query_upper: str):
        out = (C.c_int * 6)()
        r = lib().oracle_adapter_match(C.byref(self.arr[idx]), query_upper.encode(), len(query_upper), out)
        return tuple(out) if r == 1 else None


def demux_batch(rounds, seq: np.ndarray, qual: np.ndarray, offsets: np.ndarray, lengths: np.ndarray,
                n_threads: int = 1):
    """Run one or two rounds (list of (AdapterSet, revcomp)) over a batch.

    seq/qual: uint8 blobs; offsets uint64 [n]; lengths uint32 [n].
    Returns (rec_round0, rec_round1_or_None, out_seq, out_qual, out_len): the trimmed reads
    are left-aligned at the input offsets.
    """
    n = int(lengths.shape[0])
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    rec0 = np.zeros(n, dtype=MATCH_DTYPE)
    rec1 = np.zeros(n, dtype=MATCH_DTYPE)
    out_seq = np.zeros_like(seq)
    out_qual = np.zeros_like(qual)
    out_len = np.zeros(n, dtype=np.uint32)
    a0, rc0 = rounds[0]
    if len(rounds) > 1:
        a1, rc1 = rounds[1]
        p1, n1 = a1.arr, a1.n
    else:
        rc1, p1, n1 = 0, a0.arr, 0
    lib().oracle_demux_batch(len(rounds), a0.arr, a0.n, int(rc0), p1, n1, int(rc1),
                             seq.ctypes.data, qual.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, n,
                             rec0.ctypes.data, rec1.ctypes.data, out_seq.ctypes.data, out_qual.ctypes.data,
                             out_len.ctypes.data, int(n_threads))
    return rec0, (rec1 if len(rounds) > 1 else None), out_seq, out_qual, out_len


def edit_distances(seqs: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, pair_a: np.ndarray,
                   pair_b: np.ndarray, mode: str = "NW", n_threads: int = 8) -> np.ndarray:
    """edit_oracle.c: what edlib.align(shorter, longer, task='distance', mode=mode)['editDistance'] returns for
    every listed pair (amplicon_sorter.py:225-235, :838-849)."""
    L = lib()
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    out = np.zeros(pa.shape[0], dtype=np.uint32)
    L.oracle_edit_distances.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_int, C.c_void_p, C.c_int]
    L.oracle_edit_distances.restype = None
    L.oracle_edit_distances(seqs.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, pa.ctypes.data,
                            pb.ctypes.data, pa.shape[0], {"NW": 0, "HW": 1}[mode], out.ctypes.data, n_threads)
    return out
