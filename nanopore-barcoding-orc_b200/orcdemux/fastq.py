"""FASTQ(.gz) streaming in and out (what dnaio + xopen do under cutadapt).

Both directions are native (csrc/orc_io.cpp inside liborcdemux.so); this module only wraps them.
Reader: a thread inflates the file straight into a ring of pinned text buffers and finds the
record boundaries (orc_fastq_index, memchr); the raw text itself is the batch that goes to the
GPU (orc_batch raw-text layout), so no per-read copies are made on the host.  Writer: one file
per bin, created up front even if it stays empty (the reference's round-2 loop discovers its
inputs by listing them, 02_cutadapt_loop.sh:75-85); the bins' text is cut into chunks that a
thread pool deflates as gzip members, written in batch order.
"""
from __future__ import annotations

import ctypes as C
import os
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
        b = getattr(self, "bases", None)
        return b if b is not None else int(self.lengths[:self.n_reads].sum(dtype=np.uint64))


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


def _view(ptr, count, dtype):
    dt = np.dtype(dtype)
    if not ptr or count == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_uint8 * (count * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=count)


class FastqReader:
    """Iterates over TextBatch objects produced by liborcdemux.so's reader thread (orc_reader_*,
    csrc/orc_io.cpp): the input is inflated and indexed ahead of the consumer, straight into
    page-locked buffers.  The `keep` most recent batches (the one just returned included) stay
    valid; older ones go back to the reader, which owns `keep + ahead` buffers."""

    def __init__(self, path: str, max_reads: int = 1 << 18, max_bytes: int = 1 << 28, keep: int = 3,
                 ahead: int = 2, pinned: bool = True, threads: int = 0):
        self._L = _lib.load()
        self.path, self.max_reads, self.max_bytes, self.keep = path, max_reads, max_bytes, max(1, keep)
        err = C.create_string_buffer(512)
        if threads > 0:         # inflate threads for member-structured input (files orc_writer wrote)
            self._r = self._L.orc_reader_open_threads(os.fsencode(path), max_reads, max_bytes, self.keep + max(1, ahead),
                                                      int(pinned), int(threads), err, 512)
        else:
            self._r = self._L.orc_reader_open(os.fsencode(path), max_reads, max_bytes, self.keep + max(1, ahead),
                                              int(pinned), err, 512)
        if not self._r:
            raise OSError(err.value.decode(errors="replace"))
        self._held: List[int] = []

    def __iter__(self) -> Iterator[TextBatch]:
        L = self._L
        while self._r:
            while len(self._held) >= self.keep:
                L.orc_reader_release(self._r, self._held.pop(0))
            tb = _lib.TextBatchC()
            rc = L.orc_reader_next(self._r, C.byref(tb))
            if rc == 0:
                return
            if rc < 0:
                raise ValueError(L.orc_reader_error(self._r).decode(errors="replace"))
            self._held.append(int(tb.buffer))
            n = int(tb.n_reads)
            out = TextBatch(_view(tb.text, int(tb.n_bytes), np.uint8), int(tb.n_bytes), n,
                            _view(tb.offsets, n, np.uint64), _view(tb.lengths, n, np.uint32),
                            _view(tb.qual_offsets, n, np.uint64), _view(tb.name_offsets, n, np.uint64),
                            _view(tb.name_lengths, n, np.uint32))
            out.bases = int(tb.total_bases)
            yield out

    def inflate_mode(self):
        """(mode, bytes inflated by the pool, bytes inflated by the reader thread): 0 one zlib stream,
        1 member-parallel (files of orc_writer), 2 chunk-parallel (any other .gz)"""
        st = (C.c_uint64 * 2)()
        return int(self._L.orc_reader_inflate_mode(self._r, st)), int(st[0]), int(st[1])

    def close(self):
        if getattr(self, "_r", None):
            self._L.orc_reader_close(self._r)
            self._r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gunzip_file(path: str, capacity: int, threads: int = 4, chunk_bytes: int = 0) -> bytes:
    """The text of a .gz file through the chunk-parallel inflate alone (orc_gunzip_file, csrc/orc_pgz.h)."""
    L = _lib.load()
    out = np.empty(max(1, capacity), dtype=np.uint8)
    err = C.create_string_buffer(512)
    n = L.orc_gunzip_file(os.fsencode(path), int(threads), int(chunk_bytes), out.ctypes.data, int(capacity), err, 512)
    if n < 0:
        raise ValueError(err.value.decode(errors="replace"))
    return out[:n].tobytes()


class BinWriters:
    """One output file per bin name, created up front (orc_writer_*, csrc/orc_io.cpp): gzip
    members are deflated by a native thread pool while the GPU works on the next batches."""

    def __init__(self, paths: List[Optional[str]], compresslevel: int = 5, threads: int = 8, index: bool = False):
        self._L = _lib.load()
        self.paths = paths
        self.level = compresslevel
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) if p is not None else None for p in paths])
        err = C.create_string_buffer(512)
        self._w = self._L.orc_writer_open(arr, len(paths), int(compresslevel), int(threads), err, 512)
        if not self._w:
            raise OSError(err.value.decode(errors="replace"))
        if index:               # PATH.idx beside every file: (ticket, bytes) per chunk (multi-GPU merge)
            self._L.orc_writer_set_index(self._w, 1)
        self._live = {}
        self.bytes_written = [0] * len(paths)

    def write_batch(self, result, members: bool = False) -> int:
        """Queue every bin's FASTQ text of one BatchResult; returns a ticket.  The result's
        buffers must not be reused (no new submit on its slot) before wait(ticket).
        members: the result comes from an Engine(emit_gzip=True): its bins are finished gzip members, which
        .gz files receive as they are (orc_writer_write_members)."""
        off = np.ascontiguousarray(result.bin_offsets, dtype=np.uint64)
        put = self._L.orc_writer_write_members if members else self._L.orc_writer_write
        t = int(put(self._w, result.fastq.ctypes.data if result.fastq.size else None, off.ctypes.data))
        if t < 0:
            raise OSError("orc_writer_write: " + self._L.orc_writer_error(self._w).decode(errors="replace"))
        self._live[t] = (result, off)
        return t

    def wait(self, ticket: int):
        if ticket in self._live:
            rc = self._L.orc_writer_wait(self._w, ticket)
            del self._live[ticket]
            if rc:
                raise OSError("orc_writer_wait: " + self._L.orc_writer_error(self._w).decode(errors="replace"))

    def close(self):
        if getattr(self, "_w", None):
            n = np.zeros(len(self.paths), dtype=np.uint64)
            msg = self._L.orc_writer_error(self._w)
            rc = self._L.orc_writer_close(self._w, n.ctypes.data)
            self._w = None
            self._live.clear()
            self.bytes_written = [int(x) for x in n]
            if rc:
                raise OSError("writing the bins failed: " + (msg or b"").decode(errors="replace"))


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
