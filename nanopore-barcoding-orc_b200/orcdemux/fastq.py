"""FASTQ(.gz) streaming in and out (what dnaio + xopen do under cutadapt).

Reader: the file is read (inflated) straight into a ring of pinned text buffers; the record
boundaries are found by orc_fastq_index() in liborcdemux.so (C, memchr); the raw text itself
is the batch that goes to the GPU (orc_batch raw-text layout), so no per-read copies are made
on the host.  Writer: one file per bin, created up front even if it stays empty (the
reference's round-2 loop discovers its inputs by listing them, 02_cutadapt_loop.sh:75-85);
gzip members are compressed by a thread pool, one bin per task, batch order preserved.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Iterator, List, Optional

import numpy as np

from . import lib as _lib


@dataclass
class TextBatch:
    """A batch in the raw-FASTQ-text layout of orc_batch (include/orcdemux.h)."""
    text: np.ndarray            # uint8; the first n_bytes bytes are complete records
    n_bytes: int
    n_reads: int
    offsets: np.ndarray         # uint64 [>= n_reads] start of the bases
    lengths: np.ndarray         # uint32
    qual_offsets: np.ndarray    # uint64
    name_offsets: np.ndarray    # uint64
    name_lengths: np.ndarray    # uint32

    def read(self, r: int):
        t = self.text
        n0, nl = int(self.name_offsets[r]), int(self.name_lengths[r])
        o, q, L = int(self.offsets[r]), int(self.qual_offsets[r]), int(self.lengths[r])
        return (t[n0:n0 + nl].tobytes().decode(), t[o:o + L].tobytes().decode(), t[q:q + L].tobytes().decode())

    def total_bases(self) -> int:
        return int(self.lengths[:self.n_reads].sum(dtype=np.uint64))


def _pinned(n: int, dtype) -> np.ndarray:
    from .engine import pinned_empty
    return pinned_empty(n, dtype)


def index_text(text: np.ndarray, n_bytes: int, max_reads: int, final: bool, arrays=None):
    """Run orc_fastq_index over text[:n_bytes].  Returns (n_reads, consumed, arrays)."""
    L = _lib.load()
    if arrays is None:
        arrays = (np.empty(max_reads, np.uint64), np.empty(max_reads, np.uint32), np.empty(max_reads, np.uint64),
                  np.empty(max_reads, np.uint64), np.empty(max_reads, np.uint32))
    off, ln, qoff, noff, nlen = arrays
    consumed = C.c_uint64(0)
    err = C.create_string_buffer(256)
    n = L.orc_fastq_index(text.ctypes.data, n_bytes, max_reads, int(final), off.ctypes.data, ln.ctypes.data,
                          qoff.ctypes.data, noff.ctypes.data, nlen.ctypes.data, C.byref(consumed), err, 256)
    if n < 0:
        raise ValueError(err.value.decode())
    return int(n), int(consumed.value), arrays


def open_maybe_gzip(path: str):
    if path == "-":
        return sys.stdin.buffer
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rb")
    return open(path, "rb", buffering=0)


class FastqReader:
    """Iterates over TextBatch objects.  `n_buffers` batches stay valid at any time: a batch may
    be handed to the GPU while the next ones are being read (use n_buffers >= slots + 1)."""

    def __init__(self, path: str, max_reads: int = 1 << 18, max_bytes: int = 1 << 28, n_buffers: int = 4):
        self.path = path
        self.max_reads = max_reads
        self.max_bytes = max_bytes
        self.n_buffers = n_buffers
        self._bufs = [_pinned(max_bytes, np.uint8) for _ in range(n_buffers)]
        self._idx = [(_pinned(max_reads, np.uint64), _pinned(max_reads, np.uint32), _pinned(max_reads, np.uint64),
                      _pinned(max_reads, np.uint64), _pinned(max_reads, np.uint32)) for _ in range(n_buffers)]

    def __iter__(self) -> Iterator[TextBatch]:
        fh = open_maybe_gzip(self.path)
        try:
            carry = b""
            eof = False
            k = 0
            while True:
                buf = self._bufs[k % self.n_buffers]
                arrays = self._idx[k % self.n_buffers]
                fill = len(carry)
                if fill:
                    buf[:fill] = np.frombuffer(carry, dtype=np.uint8)
                mv = memoryview(buf)
                while fill < self.max_bytes and not eof:
                    got = fh.readinto(mv[fill:])
                    if not got:
                        eof = True
                        break
                    fill += got
                if fill == 0:
                    break
                n, consumed, arrays = index_text(buf, fill, self.max_reads, eof, arrays)
                if n == 0:
                    if eof:
                        break
                    raise ValueError("FASTQ record larger than the %d-byte batch buffer" % self.max_bytes)
                carry = buf[consumed:fill].tobytes()
                k += 1
                yield TextBatch(buf, consumed, n, *arrays)
                if eof and not carry:
                    break
        finally:
            if fh is not sys.stdin.buffer:
                fh.close()


class BinWriters:
    """One output file per bin name, opened (and so created) up front."""

    def __init__(self, paths: List[Optional[str]], compresslevel: int = 5, threads: int = 8):
        self.paths = paths
        self.level = compresslevel
        self._fh = []
        self._gz = []
        for p in paths:
            if p is None:
                self._fh.append(None)
                self._gz.append(None)
                continue
            self._fh.append(open(p, "wb"))
            self._gz.append(p.endswith(".gz"))
        self._pool = ThreadPoolExecutor(max_workers=max(1, threads))
        self.bytes_written = [0] * len(paths)

    def _one(self, b: int, data: bytes):
        if not data:
            return
        fh = self._fh[b]
        if self._gz[b]:
            # one gzip member per batch: a concatenation of members is a valid .gz stream
            co = zlib.compressobj(self.level, zlib.DEFLATED, 31)
            fh.write(co.compress(data) + co.flush())
        else:
            fh.write(data)
        self.bytes_written[b] += len(data)

    def write_batch(self, result):
        """Append every bin's FASTQ text of one BatchResult (bins are written in parallel,
        successive batches in order)."""
        futs = []
        for b, fh in enumerate(self._fh):
            if fh is None:
                continue
            lo, hi = int(result.bin_offsets[b]), int(result.bin_offsets[b + 1])
            if hi > lo:
                futs.append(self._pool.submit(self._one, b, result.fastq[lo:hi].tobytes()))
        for f in futs:
            f.result()

    def close(self):
        self._pool.shutdown(wait=True)
        for b, fh in enumerate(self._fh):
            if fh is None:
                continue
            if self._gz[b] and self.bytes_written[b] == 0:
                co = zlib.compressobj(self.level, zlib.DEFLATED, 31)     # an empty but valid .gz
                fh.write(co.flush())
            fh.close()


def read_adapters_fasta(path: str):
    """cutadapt parser.read_adapters_fasta: name = header.split()[0], sequence upper-cased, U -> T."""
    names, seqs = [], []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                head = line[1:].split()
                names.append(head[0] if head else "")
                seqs.append("")
            else:
                if not names:
                    raise ValueError("%s: sequence before the first FASTA header" % path)
                seqs[-1] += line
    seqs = [s.upper().replace("U", "T") for s in seqs]
    return names, seqs
