"""Host-side engine: owns one orc_ctx (one GPU) and mirrors, per round, the objects a
cutadapt invocation builds from its argv:

    Round        <- `-g file:X` / `-a file:X` + -e / -O / --no-indels / --rc
                    (parser.make_adapters_from_one_specification, adapters.FrontAdapter /
                     BackAdapter, modifiers.ReverseComplementer(AdapterCutter(times=1)))
    Engine       <- the pipeline runner: submit a batch of reads, get back per-read
                    matches, bin ids and bin-major FASTQ text (steps.Demultiplexer)

PyTorch is used for pinned host buffers only.  All arithmetic is in liborcdemux.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import lib as _lib
from .lib import ORC_BACK, ORC_FRONT


@dataclass
class Round:
    """One cutadapt invocation's adapter set (reference: 02_cutadapt_loop.sh:64-72 / :94-102)."""
    names: List[str]
    sequences: List[str]
    type: int = ORC_FRONT           # ORC_FRONT == -g, ORC_BACK == -a
    max_error_rate: float = 0.1     # -e
    min_overlap: int = 3            # -O
    indels: bool = True             # not --no-indels
    revcomp: bool = True            # --rc
    action: str = "trim"            # --action=trim | retain (keep the adapter, cut what lies beyond it)

    def __post_init__(self):
        # parser.read_adapters_fasta + adapters.SingleAdapter.__init__: upper(), U -> T
        self.sequences = [s.upper().replace("U", "T") for s in self.sequences]
        if len(self.names) != len(self.sequences):
            raise ValueError("names and sequences differ in length")
        if self.action not in ("trim", "retain"):
            raise ValueError("action must be trim or retain")


@dataclass
class BatchResult:
    n_reads: int
    matches: List[np.ndarray]       # per round, structured (lib.MATCH_DTYPE); [] unless want_matches
    bin: np.ndarray                 # int32 [n_reads], -1 = dropped
    out_len: np.ndarray             # uint32 [n_reads]
    bin_counts: np.ndarray          # uint64 [n_bins]
    bin_offsets: np.ndarray         # uint64 [n_bins + 1]
    fastq: np.ndarray               # uint8, bin-major FASTQ text

    def bin_bytes(self, b: int) -> bytes:
        return self.fastq[int(self.bin_offsets[b]):int(self.bin_offsets[b + 1])].tobytes()


class OrcError(RuntimeError):
    pass


def pinned_empty(n: int, dtype) -> np.ndarray:
    """A numpy array over page-locked memory (torch owns the allocation)."""
    import torch
    dt = np.dtype(dtype)
    # 64 bytes of slack behind the data: the kernels' 16-byte loads may over-read a buffer they read in place
    t = torch.empty(max(int(n), 1) * dt.itemsize + 64, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    # the ndarray's base chain ends in the tensor, so the page-locked block lives exactly as long as
    # the array (or any view of it) does
    return t.numpy()[:max(int(n), 1) * dt.itemsize].view(dt)[:n]


def pin_readset(rs):
    """Copy a synth.ReadSet into pinned memory (what the FASTQ reader fills in production)."""
    from .synth import ReadSet
    out = {}
    for k in ("seq", "qual", "offsets", "lengths", "names", "name_offsets"):
        src = getattr(rs, k)
        dst = pinned_empty(src.shape[0], src.dtype)
        dst[...] = src
        out[k] = dst
    return ReadSet(out["seq"], out["qual"], out["offsets"], out["lengths"], out["names"],
                   out["name_offsets"], rs.truth)


class Engine:
    """One GPU's demultiplexer.  Not thread-safe; use one Engine per GPU / process."""

    def __init__(self, rounds: Sequence[Round], device: int = 0, max_reads: int = 1 << 20,
                 max_bytes: int = 1 << 30, max_name_bytes: int = 0, n_slots: int = 2,
                 emit_fastq: bool = True, want_matches: bool = True,
                 drop_bins: Optional[np.ndarray] = None, qual_zero_copy: bool = False, emit_gzip: bool = False):
        """emit_gzip: every bin of a batch comes back as one gzip member made on the device
        (orc_params.emit_gzip): BatchResult.fastq holds the members, bin_offsets their byte ranges.
        qual_zero_copy: the qualities of a submitted batch stay in the caller's page-locked buffer and the
        emit kernel reads what it needs of them over PCIe (orc_params.qual_zero_copy; buffers from
        pinned_empty() / pin_readset() qualify: page-locked, 64 bytes of slack behind the data)."""
        if not 1 <= len(rounds) <= _lib.ORC_MAX_ROUNDS:
            raise ValueError("1 or 2 rounds")
        self._L = _lib.load()
        self.rounds = list(rounds)
        self.n_slots = n_slots
        self.want_matches = want_matches
        self.emit_fastq = emit_fastq
        p = _lib.Params()
        p.device = device
        p.n_rounds = len(rounds)
        self._keep = []
        for i, r in enumerate(rounds):
            n = len(r.sequences)
            names = (C.c_char_p * max(n, 1))(*[s.encode() for s in r.names])
            seqs = (C.c_char_p * max(n, 1))(*[s.encode() for s in r.sequences])
            self._keep += [names, seqs]
            rp = p.rounds[i]
            rp.n_adapters = n
            rp.type = r.type
            rp.names = names
            rp.sequences = seqs
            rp.max_error_rate = r.max_error_rate
            rp.min_overlap = r.min_overlap
            rp.indels = int(r.indels)
            rp.revcomp = int(r.revcomp)
            rp.action = _lib.ORC_ACTION_RETAIN if r.action == "retain" else _lib.ORC_ACTION_TRIM
        self.n_bins = 1
        for r in rounds:
            self.n_bins *= len(r.sequences) + 1
        p.max_reads = max_reads
        p.max_bytes = max_bytes
        p.max_name_bytes = max_name_bytes if max_name_bytes else 64 * max_reads
        p.n_slots = n_slots
        p.emit_fastq = int(emit_fastq)
        p.want_matches = int(want_matches)
        p.qual_zero_copy = int(qual_zero_copy)
        p.emit_gzip = int(emit_gzip)
        self.emit_gzip = bool(emit_gzip and emit_fastq)
        if drop_bins is not None:
            d = np.ascontiguousarray(drop_bins, dtype=np.uint8)
            if d.shape[0] != self.n_bins:
                raise ValueError("drop_bins must have n_bins entries")
            self._keep.append(d)
            p.drop_bins = d.ctypes.data
        err = C.create_string_buffer(512)
        self._ctx = self._L.orc_create(C.byref(p), err, 512)
        if not self._ctx:
            raise OrcError("orc_create failed: " + err.value.decode(errors="replace"))
        assert self._L.orc_n_bins(self._ctx) == self.n_bins
        self._inflight = {}

    # -- bins ------------------------------------------------------------------------
    def bin_id(self, a0: int, a1: int = -1) -> int:
        """Bin of (round-1 adapter, round-2 adapter); -1 is 'unknown'."""
        if len(self.rounds) == 1:
            return a0 + 1
        return (a0 + 1) + (len(self.rounds[0].sequences) + 1) * (a1 + 1)

    def bin_adapters(self, b: int):
        n0 = len(self.rounds[0].sequences) + 1
        if len(self.rounds) == 1:
            return (b - 1, None)
        return (b % n0 - 1, b // n0 - 1)

    # -- calls -----------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != _lib.ORC_OK:
            raise OrcError("%s failed (%d): %s" % (what, rc, self._L.orc_last_error(self._ctx).decode()))

    def _batch(self, rs) -> _lib.Batch:
        b = _lib.Batch()
        b.n_reads = rs.n_reads
        if hasattr(rs, "text"):                     # fastq.TextBatch: raw FASTQ text, one blob
            b.n_bytes = int(rs.n_bytes)
            b.seq = b.qual = rs.text.ctypes.data
            b.offsets = rs.offsets.ctypes.data
            b.lengths = rs.lengths.ctypes.data
            b.qual_offsets = rs.qual_offsets.ctypes.data
            if self.emit_fastq:
                b.names = rs.text.ctypes.data
                b.name_offsets = rs.name_offsets.ctypes.data
                b.name_lengths = rs.name_lengths.ctypes.data
            return b
        b.n_bytes = int(rs.seq.shape[0])
        b.seq = rs.seq.ctypes.data
        b.qual = rs.qual.ctypes.data
        b.offsets = rs.offsets.ctypes.data
        b.lengths = rs.lengths.ctypes.data
        if self.emit_fastq:
            b.names = rs.names.ctypes.data
            b.name_offsets = rs.name_offsets.ctypes.data
            b.name_bytes = int(rs.names.shape[0])
        return b

    def submit(self, slot: int, rs):
        """H2D + kernels + D2H of one batch, asynchronous.  `rs` must stay alive until wait()."""
        self._inflight[slot] = rs
        b = self._batch(rs)
        self._check(self._L.orc_submit(self._ctx, slot, C.byref(b)), "orc_submit")

    def upload(self, slot: int, rs):
        self._inflight[slot] = rs
        b = self._batch(rs)
        self._check(self._L.orc_upload(self._ctx, slot, C.byref(b)), "orc_upload")

    def synth(self, slot: int, seed: int, n_reads: int, len_min: int = 300, len_max: int = 900):
        """Make a shard of synthetic reads on the device (orc_synth): the slot then holds it like an upload."""
        self._inflight[slot] = None
        self._check(self._L.orc_synth(self._ctx, slot, seed, n_reads, len_min, len_max), "orc_synth")

    def export(self, slot: int):
        """The batch resident in a slot as a synth.ReadSet on the host (orc_export)."""
        from .synth import ReadSet
        n, nb, nn = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.orc_resident(self._ctx, slot, C.byref(n), C.byref(nb), C.byref(nn)), "orc_resident")
        seq = np.empty(nb.value, np.uint8)
        qual = np.empty(nb.value, np.uint8)
        off = np.empty(n.value, np.uint64)
        ln = np.empty(n.value, np.uint32)
        names = np.empty(nn.value, np.uint8)
        noff = np.empty(n.value + 1, np.uint64)
        self._check(self._L.orc_export(self._ctx, slot, seq.ctypes.data, qual.ctypes.data, off.ctypes.data,
                                       ln.ctypes.data, names.ctypes.data, noff.ctypes.data), "orc_export")
        return ReadSet(seq, qual, off, ln, names, noff, {})

    def launch(self, slot: int):
        self._check(self._L.orc_launch(self._ctx, slot), "orc_launch")

    def download(self, slot: int):
        self._check(self._L.orc_download(self._ctx, slot), "orc_download")

    def sync(self, slot: int):
        self._check(self._L.orc_sync(self._ctx, slot), "orc_sync")

    def wait(self, slot: int, copy: bool = True) -> BatchResult:
        res = _lib.Result()
        self._check(self._L.orc_wait(self._ctx, slot, C.byref(res)), "orc_wait")
        n = res.n_reads

        def view(ptr, count, dtype):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            dt = np.dtype(dtype)
            buf = (C.c_uint8 * (count * dt.itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dt, count=count)
            return a.copy() if copy else a

        matches = []
        if self.want_matches:
            for r in range(len(self.rounds)):
                matches.append(view(res.matches[r], n, _lib.MATCH_DTYPE))
        return BatchResult(
            n_reads=n, matches=matches,
            bin=view(res.bin, n, np.int32), out_len=view(res.out_len, n, np.uint32),
            bin_counts=view(res.bin_counts, self.n_bins, np.uint64),
            bin_offsets=view(res.bin_offsets, self.n_bins + 1, np.uint64),
            fastq=view(res.fastq, int(res.fastq_bytes), np.uint8))

    def run(self, rs, slot: int = 0) -> BatchResult:
        self.submit(slot, rs)
        return self.wait(slot)

    def timings(self, slot: int = 0) -> dict:
        t = _lib.Timings()
        self._check(self._L.orc_get_timings(self._ctx, slot, C.byref(t)), "orc_get_timings")
        return dict(pack_ms=t.pack_ms, trigger_ms=list(t.trigger_ms), scan_ms=list(t.scan_ms), resolve_ms=list(t.resolve_ms), bin_ms=t.bin_ms,
                    emit_ms=t.emit_ms, total_ms=t.total_ms, h2d_ms=t.h2d_ms, d2h_ms=t.d2h_ms,
                    kernel_launches=t.kernel_launches, n_tasks=list(t.n_tasks), n_candidates=list(t.n_candidates), cells=list(t.cells), cells_executed=list(t.cells_executed),
                    pack_bytes=t.pack_bytes, emit_bytes=t.emit_bytes,
                    kernel_ms=[{nm: float(t.kernel_ms[r][i]) for i, nm in enumerate(_lib.KERNEL_NAMES)}
                               for r in range(len(self.rounds))],
                    window_columns=list(t.window_columns), cells_2b=list(t.cells_2b),
                    n_pairs_2b=list(t.n_pairs_2b), n_tasks_wide=list(t.n_tasks_wide), timeline_ms=list(t.timeline_ms),
                    gzip_ms=t.gzip_ms, gzip_bytes=t.gzip_bytes)

    def timeline(self, slot: int = 0):
        """Device times (ms since the engine was made) at which the stages of the batch last waited for on the slot
        finished: upload reached, H2D done, matching done, emit done, gzip done, D2H done (orc_get_timeline)."""
        out = (C.c_float * 6)()
        self._check(self._L.orc_get_timeline(self._ctx, slot, out), "orc_get_timeline")
        return [float(x) for x in out]

    def timer_start(self, slot: int = 0):
        self._check(self._L.orc_timer_start(self._ctx, slot), "orc_timer_start")

    def timer_stop(self, slot: int = 0) -> float:
        ms = C.c_float(0.0)
        self._check(self._L.orc_timer_stop(self._ctx, slot, C.byref(ms)), "orc_timer_stop")
        return float(ms.value)

    def span_begin(self):
        """Start a device-time span over all slots (orc_span_begin)."""
        self._check(self._L.orc_span_begin(self._ctx), "orc_span_begin")

    def span_end(self) -> float:
        ms = C.c_float(0.0)
        self._check(self._L.orc_span_end(self._ctx, C.byref(ms)), "orc_span_end")
        return float(ms.value)

    def counts(self) -> np.ndarray:
        out = np.zeros(self.n_bins, dtype=np.uint64)
        self._check(self._L.orc_counts(self._ctx, out.ctypes.data), "orc_counts")
        return out

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.orc_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def m13_rounds(max_error_rate: float = 0.1, min_overlap: int = 3, revcomp: bool = True) -> List[Round]:
    """The two rounds of 02_cutadapt_loop.sh with the OrCA-seq M13 index tables."""
    from . import m13
    f, b = m13.sp5_forward(), m13.sp27_reverse_rc()
    return [Round([n for n, _ in f], [s for _, s in f], ORC_FRONT, max_error_rate, min_overlap, True, revcomp),
            Round([n for n, _ in b], [s for _, s in b], ORC_BACK, max_error_rate, min_overlap, True, revcomp)]


def measure_int32_peak(device: int = 0, mode: int = 0):
    """(32-bit integer lane-ops/s, SM clock MHz from device properties).
    mode 0: LOP3 only (ALU pipe); mode 1: LOP3 + IMAD (ALU + FMA pipes)."""
    L = _lib.load()
    clk = C.c_double(0.0)
    v = L.orc_measure_int32_peak(device, mode, C.byref(clk))
    if v <= 0:
        raise OrcError("orc_measure_int32_peak failed (no CUDA device?)")
    return v, clk.value
