"""ctypes binding of liborcdemux.so (include/orcdemux.h).

The library is the product: if it has not been built (`nanopore-barcoding-orc_b200/build.sh`
or `__graft_entry__.build()`), loading fails with an ImportError -- there is no Python or
CPU fallback for the matching path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liborcdemux.so")

ORC_MAX_ROUNDS = 2
ORC_MAX_ADAPTERS = 32
ORC_MAX_ADAPTER_LEN = 64
KERNEL_NAMES = ["sort_reads", "seed", "trigger", "sort_items", "filter", "scan", "resolve_band", "resolve_wide", "select"]
ORC_N_KERNELS = len(KERNEL_NAMES)
ORC_FRONT, ORC_BACK, ORC_PREFIX, ORC_SUFFIX = 0, 1, 2, 3
ORC_ACTION_TRIM, ORC_ACTION_RETAIN = 0, 1
ORC_OK, ORC_EINVAL, ORC_ECUDA, ORC_ECAPACITY, ORC_ESTATE = 0, -1, -2, -3, -4

# every symbol include/orcdemux.h declares
EXPORTS = ["orc_create", "orc_destroy", "orc_last_error", "orc_n_bins", "orc_submit", "orc_wait",
           "orc_upload", "orc_launch", "orc_download", "orc_sync", "orc_get_timings", "orc_timer_start", "orc_timer_stop", "orc_counts",
           "orc_fastq_index", "orc_host_alloc", "orc_host_free", "orc_measure_int32_peak", "orc_version",
           "orc_reader_open", "orc_reader_next", "orc_reader_release", "orc_reader_error", "orc_reader_close",
           "orc_writer_open", "orc_writer_write", "orc_writer_wait", "orc_writer_error", "orc_writer_close",
           "orc_edit_distances", "orc_synth", "orc_resident", "orc_export",
           "orc_reader_open_threads", "orc_writer_set_index", "orc_empty_gzip_member", "orc_span_begin", "orc_span_end", "orc_probe_hostread", "orc_writer_write_members", "orc_get_timeline", "orc_reader_inflate_mode", "orc_gunzip_file"]

MATCH_DTYPE = np.dtype([
    ("adapter", "<i4"), ("is_rc", "<i4"), ("ref_start", "<i4"), ("ref_stop", "<i4"),
    ("query_start", "<i4"), ("query_stop", "<i4"), ("score", "<i4"), ("errors", "<i4"),
])


class RoundParams(C.Structure):
    _fields_ = [("n_adapters", C.c_int32), ("type", C.c_int32),
                ("names", C.POINTER(C.c_char_p)), ("sequences", C.POINTER(C.c_char_p)),
                ("max_error_rate", C.c_double), ("min_overlap", C.c_int32),
                ("indels", C.c_int32), ("revcomp", C.c_int32), ("action", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_rounds", C.c_int32),
                ("rounds", RoundParams * ORC_MAX_ROUNDS),
                ("max_reads", C.c_uint32), ("max_bytes", C.c_uint64), ("max_name_bytes", C.c_uint64),
                ("n_slots", C.c_int32), ("emit_fastq", C.c_int32), ("want_matches", C.c_int32),
                ("drop_bins", C.c_void_p), ("qual_zero_copy", C.c_int32), ("emit_gzip", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_bytes", C.c_uint64),
                ("seq", C.c_void_p), ("qual", C.c_void_p), ("offsets", C.c_void_p), ("lengths", C.c_void_p),
                ("qual_offsets", C.c_void_p), ("names", C.c_void_p), ("name_offsets", C.c_void_p),
                ("name_lengths", C.c_void_p), ("name_bytes", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_bins", C.c_int32),
                ("matches", C.c_void_p * ORC_MAX_ROUNDS),
                ("bin", C.c_void_p), ("out_len", C.c_void_p), ("bin_counts", C.c_void_p),
                ("bin_offsets", C.c_void_p), ("fastq", C.c_void_p), ("fastq_bytes", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("pack_ms", C.c_float), ("trigger_ms", C.c_float * ORC_MAX_ROUNDS),
                ("scan_ms", C.c_float * ORC_MAX_ROUNDS),
                ("resolve_ms", C.c_float * ORC_MAX_ROUNDS), ("bin_ms", C.c_float), ("emit_ms", C.c_float),
                ("total_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("n_tasks", C.c_uint32 * ORC_MAX_ROUNDS),
                ("n_candidates", C.c_uint32 * ORC_MAX_ROUNDS),
                ("cells", C.c_uint64 * ORC_MAX_ROUNDS), ("cells_executed", C.c_uint64 * ORC_MAX_ROUNDS), ("pack_bytes", C.c_uint64), ("emit_bytes", C.c_uint64),
                ("kernel_ms", (C.c_float * ORC_N_KERNELS) * ORC_MAX_ROUNDS),
                ("window_columns", C.c_uint64 * ORC_MAX_ROUNDS), ("cells_2b", C.c_uint64 * ORC_MAX_ROUNDS),
                ("n_pairs_2b", C.c_uint32 * ORC_MAX_ROUNDS), ("n_tasks_wide", C.c_uint32 * ORC_MAX_ROUNDS),
                ("timeline_ms", C.c_float * 5), ("gzip_ms", C.c_float), ("gzip_bytes", C.c_uint64)]


class TextBatchC(C.Structure):
    _fields_ = [("text", C.c_void_p), ("n_bytes", C.c_uint64), ("n_reads", C.c_uint32), ("buffer", C.c_int32),
                ("offsets", C.c_void_p), ("lengths", C.c_void_p), ("qual_offsets", C.c_void_p),
                ("name_offsets", C.c_void_p), ("name_lengths", C.c_void_p), ("total_bases", C.c_uint64)]


_lib = None


def load():
    """Load liborcdemux.so and declare prototypes.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "liborcdemux.so is not built (%s). Run nanopore-barcoding-orc_b200/build.sh or "
            "__graft_entry__.build(); orcdemux has no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    L.orc_create.argtypes = [C.POINTER(Params), C.c_char_p, C.c_size_t]
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_destroy.restype = None
    L.orc_last_error.argtypes = [C.c_void_p]
    L.orc_last_error.restype = C.c_char_p
    L.orc_n_bins.argtypes = [C.c_void_p]
    L.orc_n_bins.restype = C.c_int
    for name in ("orc_submit", "orc_upload"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.POINTER(Batch)]
        getattr(L, name).restype = C.c_int
    for name in ("orc_launch", "orc_download", "orc_sync"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        getattr(L, name).restype = C.c_int
    L.orc_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(Result)]
    L.orc_wait.restype = C.c_int
    L.orc_get_timings.argtypes = [C.c_void_p, C.c_int, C.POINTER(Timings)]
    L.orc_get_timings.restype = C.c_int
    L.orc_get_timeline.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.orc_get_timeline.restype = C.c_int
    L.orc_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_counts.restype = C.c_int
    L.orc_fastq_index.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
    L.orc_fastq_index.restype = C.c_int64
    L.orc_host_alloc.argtypes = [C.c_size_t]
    L.orc_host_alloc.restype = C.c_void_p
    L.orc_host_free.argtypes = [C.c_void_p]
    L.orc_host_free.restype = None
    L.orc_timer_start.argtypes = [C.c_void_p, C.c_int]
    L.orc_timer_start.restype = C.c_int
    L.orc_timer_stop.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.orc_timer_stop.restype = C.c_int
    L.orc_span_begin.argtypes = [C.c_void_p]
    L.orc_span_begin.restype = C.c_int
    L.orc_span_end.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.orc_span_end.restype = C.c_int
    L.orc_measure_int32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.orc_measure_int32_peak.restype = C.c_double
    L.orc_probe_hostread.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64]
    L.orc_probe_hostread.restype = C.c_double
    L.orc_reader_open.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_reader_open.restype = C.c_void_p
    L.orc_reader_open_threads.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_reader_open_threads.restype = C.c_void_p
    L.orc_writer_set_index.argtypes = [C.c_void_p, C.c_int]
    L.orc_writer_set_index.restype = C.c_int
    L.orc_empty_gzip_member.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.orc_empty_gzip_member.restype = C.c_size_t
    L.orc_reader_next.argtypes = [C.c_void_p, C.POINTER(TextBatchC)]
    L.orc_reader_next.restype = C.c_int
    L.orc_reader_release.argtypes = [C.c_void_p, C.c_int]
    L.orc_reader_release.restype = C.c_int
    L.orc_reader_inflate_mode.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_reader_inflate_mode.restype = C.c_int
    L.orc_gunzip_file.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_size_t]
    L.orc_gunzip_file.restype = C.c_int64
    L.orc_reader_error.argtypes = [C.c_void_p]
    L.orc_reader_error.restype = C.c_char_p
    L.orc_reader_close.argtypes = [C.c_void_p]
    L.orc_reader_close.restype = None
    L.orc_writer_open.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_writer_open.restype = C.c_void_p
    L.orc_writer_write.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_writer_write.restype = C.c_int64
    L.orc_writer_write_members.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_writer_write_members.restype = C.c_int64
    L.orc_writer_wait.argtypes = [C.c_void_p, C.c_int64]
    L.orc_writer_wait.restype = C.c_int
    L.orc_writer_error.argtypes = [C.c_void_p]
    L.orc_writer_error.restype = C.c_char_p
    L.orc_writer_close.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_writer_close.restype = C.c_int
    L.orc_edit_distances.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_char_p, C.c_size_t]
    L.orc_edit_distances.restype = C.c_int
    L.orc_synth.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    L.orc_synth.restype = C.c_int
    L.orc_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_resident.restype = C.c_int
    L.orc_export.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.orc_export.restype = C.c_int
    L.orc_version.argtypes = []
    L.orc_version.restype = C.c_char_p
    _lib = L
    return L
