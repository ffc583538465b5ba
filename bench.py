#!/usr/bin/env python
"""bench.py -- reads/s of the two-round SP5 x SP27 demultiplex (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (B200, liborcdemux.so)
    python bench.py --impl reference [...]                       the reference's CPU path

A "step" is one pass of the whole hot path (pack, round-1 seed/trigger/filter/scan/resolve/select,
round 2 the same, bin partition, FASTQ emit) over one batch of 1 Mi synthetic reads.

N = 1   BASELINE configs[1]: 1 Mi synthetic COI-length reads (300-900 nt, numpy generator, seed 1002), full
        two-round demux + trim; the line also carries short runs of configs[2] and configs[3] (`extra_configs`).
N > 1   BASELINE configs[4] (torchrun, one rank per GPU): every step of every rank is a DISTINCT shard of 1 Mi
        reads of the same read model, made on the GPU by orc_synth() with seed (1005 << 32) + shard, shard =
        rank + world * step, until 100 M reads exist in total (later steps revisit the rank's shards).  All
        shards are resident in HBM before the timed region.  No data-path collective; the 169 per-bin
        counters are all-reduced at the end, and every rank checks the bins of a sampled part of one of its
        shards against the CPU oracle.

`value`   reads/s with the batches resident in HBM (kernels only), CUDA events on the library's stream, max
          over ranks.
`e2e`     the same metric through the C ABI with host buffers: every step copies its inputs from pinned host
          memory and its results (the FASTQ text of the 96 bins the reference script keeps, bin ids, trimmed
          lengths, counters) back.  `e2e.link` says how much of the host<->device ceiling measured in the same
          run (all ranks copying at once) that is.
`roofline`  the kernel that takes the largest share of the step, against its own roof; `roofline_kernels`
          lists every kernel.  DP kernels: DP cells (or seed probes) really executed x ALU-pipe instructions
          per cell column (counted in the SASS, profiles/README.md) against the LOP3 issue rate measured in
          this run by orc_measure_int32_peak().  The exact filters skip most of the algorithmic cells
          (2 * 12 * m * n per read and round); that factor is `algorithmic_speedup`, not a
          roofline fraction.  Copy kernels: algorithmic bytes against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / `--impl reference`: real cutadapt if one is on PATH / importable / under baseline/_ref
          (kind "reference"), else the oracle's C restatement of it (kind "port"), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "nanopore-barcoding-orc_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "reads/sec demuxed (2-round SP5xSP27)"
METRIC_R1 = "reads/sec demuxed (round 1, SP5 5' indices)"        # --config 1 only
UNIT = "reads/s"
TARGET_READS = 100_000_000          # BASELINE configs[4]

# ALU-pipe instructions (LOP3 / IADD3 / SHF / PRMT / ISETP ...; IMAD runs on the FMA pipe) per column of the
# inner loops, counted in the SASS of this build (profiles/README.md, "SASS counts"), and the rows a column
# step updates.  peak cells/s of a kernel = measured LOP3 lane-op/s / ALU instructions per column x rows.
ALU_PER_COLUMN = {"filter": 12.0, "scan": 24.0, "resolve_band": 20.0, "seed": 2.25}
ROWS = {"filter": 32.0, "scan": None, "resolve_band": 32.0, "seed": 1.0}      # scan: the adapter length m


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1 << 20)
    ap.add_argument("--len-min", type=int, default=300)
    ap.add_argument("--len-max", type=int, default=900)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short configs[2] / configs[3] runs")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs[]: 1 = round 1 only (SP5 5' demux) on 100 k reads, the reference's "
                         "own CPU-runnable case; 2 = 1 Mi COI reads two-round (default at N = 1, the metric's config), "
                         "3 = rRNA-cistron reads 1-3.5 kb two-round, 4 = anchored --no-indels Hamming path "
                         "(24 M13 variable indices) on reads with the bare index at offset 0, 5 = distinct shards "
                         "made on the GPU (the default under torchrun)")
    ap.add_argument("--sub-batches", type=int, default=8)
    ap.add_argument("--qual-copy", action="store_true",
                    help="e2e leg: copy the quality strings to the device whole instead of letting the emit kernel "
                         "read what it needs of them from the pinned host buffer (orc_params.qual_zero_copy)")
    ap.add_argument("--e2e-text", action="store_true",
                    help="e2e leg: the bins come back as plain FASTQ text instead of gzip members coded on the device "
                         "(orc_params.emit_gzip); without this flag the plain-text variant is reported beside it as e2e_text")
    ap.add_argument("--resident", type=int, default=2,
                    help="batches resident in HBM whose steps are in flight together on their own streams (device-resident leg)")
    ap.add_argument("--oracle-sample", type=int, default=32768, help="reads of one shard per rank checked against the oracle (config 5)")
    a = ap.parse_args()
    if a.config == 1 and a.reads == 1 << 20:
        a.reads = 100000
    if a.config == 3 and a.len_min == 300 and a.len_max == 900:
        a.len_min, a.len_max = 1000, 3500
        if a.reads == 1 << 20:
            a.reads = 1 << 18
    return a


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled by
    a thread every millisecond (the timed region is tens of milliseconds, far shorter than one
    `nvidia-smi -lms` period); the device is found by UUID so CUDA_VISIBLE_DEVICES does not matter."""
    NAMES = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
             ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
             ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")]

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.h = None
        self.sm, self.bits, self.power = [], 0, []
        self._stop = threading.Event()
        self.t = None
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = nv
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:          # no NVML: say so in the line instead of inventing clocks
            self.h = None
            self.why = "NVML unavailable: %s" % e

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.why]}
        self._stop.set()
        self.t.join(timeout=2)
        reasons = [nm for nm, attr in self.NAMES if self.bits & int(getattr(self.nv, attr))]
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU arms
def find_real_cutadapt():
    """(kind, handle): a real cutadapt executable (not this repo's shim) or an importable package; else (None, None)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    shim_dir = os.path.realpath(os.path.join(PKG, "bin"))
    for d in os.environ.get("PATH", "").split(os.pathsep) + [os.path.join(ref, "bin")]:
        if not d or os.path.realpath(d) == shim_dir:
            continue
        p = os.path.join(d, "cutadapt")
        if os.path.isfile(p) and os.access(p, os.X_OK):
            try:
                with open(p, "rb") as fh:
                    if b"orcdemux" in fh.read(4096):
                        continue
            except OSError:
                continue
            return "exe", [p]
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import importlib.util
        if importlib.util.find_spec("cutadapt") is not None:
            return "module", [sys.executable, "-m", "cutadapt"]
    except Exception:
        pass
    return None, None


def real_cutadapt_arm(cmd, rs, sub_n, threads, steps, warmup, n_rounds):
    """Time the reference script's own command lines (02_cutadapt_loop.sh:64-72, 91-103) with a real cutadapt on
    the first sub_n reads, uncompressed FASTQ on /dev/shm (the I/O-light variant of SURVEY 8d)."""
    from orcdemux import m13, synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="orc_ref_", dir=base)
    try:
        recs = [rs.read(r) for r in range(sub_n)]
        sub = synth.from_records(recs)
        inp = os.path.join(tmp, "pychopped_ds.fastq")
        with open(inp, "wb") as fh:
            fh.write(sub.to_fastq_bytes())
        fwd, rev = os.path.join(tmp, "fwd.fa"), os.path.join(tmp, "rev.fa")
        with open(fwd, "w") as fh:
            fh.write("".join(">%s\n%s\n" % x for x in m13.sp5_forward()))
        with open(rev, "w") as fh:
            fh.write("".join(">%s\n%s\n" % x for x in m13.sp27_reverse_rc()))
        times = []
        for it in range(warmup + steps):
            for d in ("SP5", "SP27"):
                shutil.rmtree(os.path.join(tmp, d), ignore_errors=True)
                os.makedirs(os.path.join(tmp, d))
            t0 = time.perf_counter()
            subprocess.run(cmd + ["--action=trim", "-e", "0.1", "-j", str(threads), "--rc", "-g", "file:" + fwd,
                                  "-o", os.path.join(tmp, "SP5", "{name}_ds.fastq"), inp,
                                  "--json=" + os.path.join(tmp, "SP5", "r1.json")], check=True,
                           stdout=subprocess.DEVNULL)
            if n_rounds > 1:
                for nm, _ in m13.sp5_forward():
                    subprocess.run(cmd + ["--action=trim", "-e", "0.1", "-j", str(threads), "--rc", "-a", "file:" + rev,
                                          "-o", os.path.join(tmp, "SP27", "{name}_%s_ds.fastq" % nm),
                                          os.path.join(tmp, "SP5", nm + "_ds.fastq"),
                                          "--json=" + os.path.join(tmp, "SP27", nm + ".json")], check=True,
                                   stdout=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        return times
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_arm(rs, n_sample, threads, steps, warmup, n_rounds=2):
    """-> (reads, times, kind, how): real cutadapt when there is one, else the oracle's C restatement, on the
    first n_sample reads."""
    sub_n = min(n_sample, rs.n_reads)
    kind, cmd = find_real_cutadapt()
    if kind is not None:
        try:
            times = real_cutadapt_arm(cmd, rs, sub_n, threads, steps, warmup, n_rounds)
            return sub_n, times, "reference", "real cutadapt (%s) -j %d, the script's own command lines, uncompressed FASTQ on /dev/shm" % (" ".join(cmd), threads)
        except Exception as e:          # a broken install must not take the bench line down
            sys.stderr.write("real cutadapt failed (%s): falling back to the port\n" % e)
    import oracle
    from orcdemux import m13
    oracle.build()
    end = int(rs.offsets[sub_n - 1] + rs.lengths[sub_n - 1]) if sub_n else 0
    seq, qual = rs.seq[:end], rs.qual[:end]
    off, ln = rs.offsets[:sub_n], rs.lengths[:sub_n]
    sets = [(oracle.AdapterSet([q for _, q in m13.sp5_forward()], oracle.FRONT, 0.1, 3), 1),
            (oracle.AdapterSet([q for _, q in m13.sp27_reverse_rc()], oracle.BACK, 0.1, 3), 1)][:n_rounds]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.demux_batch(sets, seq, qual, off, ln, n_threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sub_n, times, "port", ("oracle/cutadapt_oracle.c (restated cutadapt 4.9: Ukkonen-banded DP, no k-mer prefilter, "
                                  "gcc -O3 -march=x86-64-v3, scratch buffers hoisted, %d pthreads); cutadapt itself is not "
                                  "vendored in the reference and not installed here" % threads)


# --------------------------------------------------------------------------- pieces of our arm
def scripts_final_tree_drop(n5=12, n27=12):
    """drop_bins of the tree 02_cutadapt_loop.sh leaves (02:107-119): no 'unknown' of either round, no SP27_009..012."""
    drop = np.zeros((n5 + 1) * (n27 + 1), dtype=np.uint8)
    for b in range(drop.shape[0]):
        i5, i27 = b % (n5 + 1) - 1, b // (n5 + 1) - 1
        if i5 < 0 or i27 < 0 or i27 >= 8:
            drop[b] = 1
    return drop


def split_subbatches(E, synth, rs, n_sub):
    """A ReadSet cut into n_sub contiguous sub-batches in pinned memory (what a FASTQ reader hands over)."""
    per = (rs.n_reads + n_sub - 1) // n_sub
    subs = []
    for i in range(n_sub):
        lo, hi = i * per, min(rs.n_reads, (i + 1) * per)
        if lo >= hi:
            break
        b0 = int(rs.offsets[lo])
        b1 = int(rs.offsets[hi - 1]) + int(rs.lengths[hi - 1])
        n0, n1 = int(rs.name_offsets[lo]), int(rs.name_offsets[hi])
        off = E.pinned_empty(hi - lo, np.uint64)
        off[...] = rs.offsets[lo:hi] - np.uint64(b0)
        noff = E.pinned_empty(hi - lo + 1, np.uint64)
        noff[...] = rs.name_offsets[lo:hi + 1] - np.uint64(n0)
        subs.append(synth.ReadSet(rs.seq[b0:b1], rs.qual[b0:b1], off, rs.lengths[lo:hi], rs.names[n0:n1], noff, {}))
    return subs, per


def run_e2e(E, rounds, device, step_batches, steps, barrier, drop, want_matches, S=4, zero_copy=True, gzip=False):
    """`steps` steps through orc_submit/orc_wait with host buffers; step k streams the sub-batches of
    step_batches[k % len(step_batches)] over S slots.  -> (seconds, h2d bytes per step, d2h bytes per step,
    reads, cumulative counts, bytes of h2d that the emit kernel read in place).

    zero_copy: orc_params.qual_zero_copy -- the quality strings are not copied to the device; the emit kernel
    reads the trimmed qualities of the reads whose bin is kept straight from the pinned host buffer.  Those
    bytes cross PCIe all the same and are counted in h2d (from the results of the untimed warm-up step)."""
    per = max(x.n_reads for sb in step_batches for x in sb)
    eng = E.Engine(rounds, device=device, max_reads=per,
                   max_bytes=max(int(x.seq.shape[0]) for sb in step_batches for x in sb) + 64,
                   max_name_bytes=max(int(x.names.shape[0]) for sb in step_batches for x in sb) + 64, n_slots=S,
                   emit_fastq=True, want_matches=want_matches, drop_bins=drop, qual_zero_copy=zero_copy, emit_gzip=gzip)
    h2d = sum(int(x.seq.nbytes + (0 if zero_copy else x.qual.nbytes) + x.offsets.nbytes + x.lengths.nbytes +
                  x.names.nbytes + x.name_offsets.nbytes) for x in step_batches[0])
    state = {"inflight": [], "k": 0, "reads": 0, "d2h": 0, "in_place": 0, "warm": True}

    def take(slot):
        r = eng.wait(slot, copy=False)
        state["reads"] += r.n_reads
        if state["warm"] and zero_copy:         # quality bytes the emit kernel fetched from host memory
            state["in_place"] += int(r.out_len[r.bin >= 0].sum())
        state["d2h"] += int(r.fastq.nbytes + r.bin.nbytes + r.out_len.nbytes + r.bin_counts.nbytes +
                            r.bin_offsets.nbytes + sum(m.nbytes for m in r.matches))

    def pump(sub):
        if len(state["inflight"]) == S:
            take(state["inflight"].pop(0))
        slot = state["k"] % S
        eng.submit(slot, sub)
        state["inflight"].append(slot)
        state["k"] += 1

    def drain():
        while state["inflight"]:
            take(state["inflight"].pop(0))

    for sub in step_batches[0]:                 # warm the copy paths: one untimed step
        pump(sub)
    drain()
    state.update(reads=0, d2h=0, warm=False)
    barrier()
    w0 = time.perf_counter()
    for k in range(steps):
        for sub in step_batches[k % len(step_batches)]:
            pump(sub)
    drain()
    barrier()
    secs = time.perf_counter() - w0
    counts = eng.counts().astype(np.int64)
    eng.close()
    return secs, h2d + state["in_place"], state["d2h"] // max(steps, 1), state["reads"], counts, state["in_place"]


def measure_link(torch, barrier, n_bytes=1 << 29):
    """Host<->device copy ceiling of THIS rank while every rank does the same (pinned 512 MiB blocks):
    H2D alone, D2H alone, both directions at once (GB/s per direction)."""
    a = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    b = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for name in ("h2d", "d2h", "both"):
        best = None
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d.copy_(a, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    b.copy_(d2, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best):
                best = dt
        out[name + "_gbs"] = n_bytes / best / 1e9
    return out


def kernel_table(t, rounds, alu_peak, hbm, hbm_how, m_rows, seed_columns):
    """roofline_kernels: every kernel of the last step with its time, share and roof."""
    rows = []
    total = t["total_ms"]
    n_rounds = len(rounds)

    def add(name, ms, bound=None, achieved=None, peak=None, unit=None, how=None):
        if ms <= 0:
            return
        r = {"kernel": name, "ms": ms, "share": ms / total if total else None, "bound": bound,
             "achieved": achieved, "peak": peak, "unit": unit,
             "frac": (achieved / peak) if (achieved is not None and peak) else None}
        if how:
            r["how"] = how
        rows.append(r)

    def alu(name, key, ms, cols_times_rows, rows_per_col):
        peak = alu_peak / ALU_PER_COLUMN[key] * rows_per_col / 1e9
        add(name, ms, "int32_alu", cols_times_rows / (ms * 1e-3) / 1e9 if ms > 0 else None, peak, "GCUPS",
            "executed cells / time vs LOP3 rate / %.4g ALU-pipe instr per column x %.4g rows" % (ALU_PER_COLUMN[key], rows_per_col))

    add("pack_kernel", t["pack_ms"], "hbm", t["pack_bytes"] / (t["pack_ms"] * 1e-3) / 1e9 if t["pack_ms"] > 0 else None,
        hbm, "GB/s", "1.5 B per input byte; " + hbm_how)
    for r in range(n_rounds):
        k = t["kernel_ms"][r]
        tag = " r%d" % (r + 1)
        add("bucket_scatter_kernel<reads>" + tag, k["sort_reads"], "hbm")
        if k["seed"] > 0:
            alu("seed_kernel" + tag, "seed", k["seed"], float(seed_columns[r]), 1.0)
        add("trigger_kernel" + tag, k["trigger"], "int32_alu")
        add("bucket_scatter_kernel<items>" + tag, k["sort_items"], "hbm")
        if k["filter"] > 0:
            # every adapter's block rows over the window columns (pairs that pass leave early: an upper bound)
            block_rows = float(sum(min(32, len(q)) for q in rounds[r].sequences))
            alu("filter_kernel" + tag, "filter", k["filter"], block_rows * float(t["window_columns"][r]), 32.0)
        if k["scan"] > 0 and t["cells_2b"][r]:
            alu("scan_kernel" + tag, "scan", k["scan"], float(t["cells_2b"][r]), m_rows[r])
        add("resolve_band_kernel" + tag, k["resolve_band"], "int32_alu")
        add("resolve_kernel (wide)" + tag, k["resolve_wide"], "latency")
        add("select_kernel" + tag, k["select"], "hbm")
    add("bin_count/scan/offsets/place", t["bin_ms"], "hbm")
    add("emit_kernel", t["emit_ms"], "hbm", t["emit_bytes"] / (t["emit_ms"] * 1e-3) / 1e9 if t["emit_ms"] > 0 else None,
        hbm, "GB/s", "every FASTQ byte read once and written once; " + hbm_how)
    return rows


def device_resident(E, eng, n_slots, steps, warmup, barrier, sampler):
    """K launches over the resident batches, step k on slot k % n_slots (every slot has its own stream, so the
    kernels of consecutive steps overlap like they do in the submit/wait pipeline); device time of the whole
    span from CUDA events that bracket all streams (orc_span_begin / orc_span_end)."""
    n = n_slots
    for k in range(max(warmup, n)):
        eng.launch(k % n)
    for s in range(n):
        eng.sync(s)
    barrier()
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    eng.span_begin()
    for k in range(steps):
        eng.launch(k % n)
    dev_ms = eng.span_end()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop() if sampler else None
    # the stage split comes from one launch with nothing else in flight
    eng.launch(0)
    eng.sync(0)
    return dev_ms, wall_ms, clocks


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: everything else (NCCL's version banner, library
    # chatter) goes to stderr until the line is printed
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    ncpu = os.cpu_count() or 1
    if world > 1 and args.config == 2:
        args.config = 5

    from orcdemux import synth

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = ncpu
        one_round = args.config == 1
        n_sample = args.cpu_sample or (100000 if one_round else 16384)
        seed = 1001 if one_round else (1005 << 32 if args.config == 5 else 1002)
        rs = synth.generate(n_sample, args.len_min, args.len_max, seed=seed, workers=min(8, ncpu))
        warm = min(args.warmup, 1)
        sub_n, times, kind, how = cpu_arm(rs, n_sample, threads, args.steps, warm, 1 if one_round else 2)
        ms = 1e3 * float(np.mean(times))
        val = sub_n / float(np.mean(times))
        line = {
            "impl": "reference", "metric": METRIC_R1 if one_round else METRIC, "value": val, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": ("configs[0]: round-1 SP5 5' demux (-g file:M13_amplicon_indices_forward.fa -e 0.1 --rc), "
                                    "%d synthetic reads %d-%d nt, seed 1001" % (sub_n, args.len_min, args.len_max))
                       if one_round else
                       "%s: two-round SP5->SP27 demux + trim, synthetic COI reads %d-%d nt; each step = %d reads "
                       "of the workload's read model (numpy generator, seed %d)"
                       % ("configs[4]" if args.config == 5 else "configs[1]", args.len_min, args.len_max, sub_n, seed),
                       "reads_per_step": sub_n, "cpu_arm": how},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d reads of the workload, %d timed passes; %s" % (sub_n, args.steps, how)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        if one_round:       # algorithmic cells of round 1: 2 orientations x 12 adapters x 59 rows x bases
            line["gcups"] = 2 * 12 * 59 * float(rs.lengths[:sub_n].sum()) / float(np.mean(times)) / 1e9
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from orcdemux import engine as E

    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: bench.py has no CPU fallback for the product arm"})
        return 2
    torch.cuda.set_device(local_rank)
    # pinned buffers should live on the GPU's own NUMA node: bind this rank to the CPUs NVML names
    # for the device before anything page-locked is allocated (undone for the CPU baseline)
    all_cpus = os.sched_getaffinity(0)
    numa = "unbound"
    if os.environ.get("ORC_BENCH_AFFINITY", "1") == "1":
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)).encode())
            words = nv.nvmlDeviceGetCpuAffinity(h, (max(all_cpus) // 64) + 1)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & all_cpus
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = "%d of %d cpus" % (len(cpus), len(all_cpus))
        except Exception as e:
            numa = "unbound (%s)" % type(e).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_config(config, reads, len_min, len_max, steps, warmup, full):
        """One workload: device-resident value, e2e, stage split.  `full` adds clocks, the kernel table inputs."""
        from orcdemux import m13
        from orcdemux.lib import ORC_PREFIX
        out = {}
        t0 = time.perf_counter()
        if config == 4:
            var = m13.variable_all()
            rounds = [E.Round([n for n, _ in var], [q for _, q in var], ORC_PREFIX, 0.1, 3, False, True)]
        elif config == 1:
            rounds = E.m13_rounds()[:1]
        else:
            rounds = E.m13_rounds()
        drop = scripts_final_tree_drop() if len(rounds) == 2 else None
        oracle_check = None
        if config == 5:
            # distinct shards made on the GPU, all resident before the timed region
            per_rank = max(1, min(steps, -(-TARGET_READS // (reads * world))))
            shard_ids = [rank + world * k for k in range(per_rank)]
            eng = E.Engine(rounds, device=local_rank, max_reads=reads, max_bytes=int(reads * (len_max + 72)),
                           max_name_bytes=24 * reads, n_slots=per_rank, emit_fastq=True, want_matches=False,
                           drop_bins=drop)
            for s, sid in enumerate(shard_ids):
                eng.synth(s, (1005 << 32) + sid, reads, len_min, len_max)
            for s in range(per_rank):
                eng.sync(s)
            out["gen_s"] = time.perf_counter() - t0
            rs_list = [None] * per_rank
            seed_desc = "(1005 << 32) + shard, shard = rank + %d * step, %d shards per rank" % (world, per_rank)
            n_bytes = None
        else:
            seed = {1: 1001, 2: 1002, 3: 1003, 4: 1004}[config]
            workers = max(1, min(16, ncpu // max(world, 1)))
            rs = synth.generate(reads, len_min, len_max, seed=seed, workers=workers, anchored=(config == 4))
            out["gen_s"] = time.perf_counter() - t0
            rs = E.pin_readset(rs)
            n_bytes = int(rs.seq.shape[0])
            n_res = max(1, args.resident)
            eng = E.Engine(rounds, device=local_rank, max_reads=rs.n_reads, max_bytes=n_bytes,
                           max_name_bytes=int(rs.names.shape[0]) + 64, n_slots=n_res, emit_fastq=True, want_matches=False,
                           drop_bins=drop)
            for sl in range(n_res):                 # the same batch resident in every slot
                eng.upload(sl, rs)
                eng.sync(sl)
            rs_list = [rs] * n_res
            seed_desc = str(seed)
        sampler = ClockSampler(local_rank) if full else None
        dev_ms, wall_ms, clocks = device_resident(E, eng, len(rs_list), steps, warmup, barrier, sampler)
        t = eng.timings(0)
        tm = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        max_ms = float(tm.item())
        out.update(value=(reads * world * steps) / (max_ms * 1e-3), ms_per_step=max_ms / steps,
                   wall_ms_per_step=wall_ms / steps, clocks=clocks, timings=t, seed=seed_desc, n_bytes=n_bytes,
                   launches=int(t["kernel_launches"]) * steps, rounds=rounds, n_slots=len(rs_list))
        # ---- sampled oracle check of one shard per rank (config 5) and host copies for the e2e leg
        host_batches = None
        if config == 5:
            import oracle
            k = len(rs_list) // 2
            eng.launch(k)
            eng.download(k)
            res = eng.wait(k)
            shard = eng.export(k)               # the generated reads, as the host sees them
            ns = min(args.oracle_sample, shard.n_reads)
            end = int(shard.offsets[ns - 1] + shard.lengths[ns - 1])
            sets = [(oracle.AdapterSet(r.sequences, oracle.FRONT if i == 0 else oracle.BACK, 0.1, 3), 1) for i, r in enumerate(rounds)]
            os.sched_setaffinity(0, all_cpus)
            t1 = time.perf_counter()
            rec0, rec1, _, _, olen = oracle.demux_batch(sets, shard.seq[:end], shard.qual[:end], shard.offsets[:ns],
                                                        shard.lengths[:ns], n_threads=max(1, ncpu // world))
            exp_bin = (rec0["adapter"] + 1) + 13 * (rec1["adapter"] + 1)
            exp_bin = np.where(drop[exp_bin] != 0, -1, exp_bin).astype(np.int32)
            ok = bool(np.array_equal(res.bin[:ns], exp_bin) and np.array_equal(res.out_len[:ns], olen))
            exp_counts = np.bincount(exp_bin[exp_bin >= 0], minlength=169)
            got_counts = np.bincount(res.bin[:ns][res.bin[:ns] >= 0], minlength=169)
            oracle_check = {"shard": shard_ids[k], "reads": ns, "bins_equal": ok,
                            "counts_equal": bool(np.array_equal(exp_counts, got_counts)),
                            "seconds": time.perf_counter() - t1,
                            "binned_of_sample": int(got_counts.sum())}
            if not (ok and oracle_check["counts_equal"]):
                raise SystemExit("rank %d: shard %d differs from the oracle" % (rank, shard_ids[k]))
            host_batches = [E.pin_readset(shard)]
            if len(rs_list) > 1:
                eng.launch(0); eng.download(0); eng.wait(0)
                host_batches.append(E.pin_readset(eng.export(0)))
        else:
            host_batches = rs_list[:1]
        eng.close()
        out["oracle_check"] = oracle_check
        # ---- end to end through the C ABI with host buffers
        if not args.no_e2e:
            step_batches = [split_subbatches(E, synth, hb, max(1, args.sub_batches))[0] for hb in host_batches]
            e_steps = steps if full else max(2, min(steps, 5))
            zc = not args.qual_copy

            def e2e_leg(gz):
                secs, h2d, d2h, n_done, counts, in_place = run_e2e(E, rounds, local_rank, step_batches, e_steps, barrier, drop,
                                                                   want_matches=False, zero_copy=zc, gzip=gz)
                te = torch.tensor([secs], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(te, op=dist.ReduceOp.MAX)
                assert n_done == e_steps * reads, (n_done, e_steps, reads)
                return counts, {
                    "value": (reads * world * e_steps) / float(te.item()), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                    "ms_per_step": 1e3 * float(te.item()) / e_steps,
                    "output": "gzip members" if gz else "FASTQ text",
                    "note": "each step streamed as %d sub-batches over 4 slots/streams (copies overlap kernels); "
                            "pipeline fill and drain are inside the timed region; output = the bins the reference "
                            "script keeps (02:107-119: no unknown, no SP27_009..012) + bin id and trimmed length per "
                            "read; match records off" % len(step_batches[0]) +
                            ("; every bin of a sub-batch comes back as one gzip member coded on the device (orc_params."
                             "emit_gzip, csrc/orc_gz.cuh: what 02:64-72 / 02:94-102 leave on disk are .fastq.gz files), "
                             "only the compressed bytes cross PCIe" if gz else "; the bins come back as FASTQ text") +
                            ("; the quality strings are NOT copied to the device: the emit kernel reads the trimmed "
                             "qualities of the kept reads in place from the pinned host buffer (orc_params."
                             "qual_zero_copy), %d of the h2d bytes per step" % in_place if zc else
                             "; quality strings copied to the device whole (--qual-copy)"),
                    "h2d_in_place_bytes_per_step": in_place}

            counts, out["e2e"] = e2e_leg(not args.e2e_text)
            if not args.e2e_text and full:
                counts_t, out["e2e_text"] = e2e_leg(False)
                assert np.array_equal(counts, counts_t)
            if full and rank == 0 and not args.e2e_text and len(rounds) == 2:
                # device time of the gzip stage on one whole resident batch (it is not part of `value`)
                hb = host_batches[0]
                with E.Engine(rounds, device=local_rank, max_reads=hb.n_reads, max_bytes=int(hb.seq.shape[0]) + 64,
                              max_name_bytes=int(hb.names.shape[0]) + 64, n_slots=1, emit_fastq=True,
                              want_matches=False, drop_bins=drop, emit_gzip=True) as gz:
                    rg = gz.run(hb)
                    # every 8th bin of the whole batch inflated by zlib and compared with the text path's bytes
                    import zlib
                    with E.Engine(rounds, device=local_rank, max_reads=hb.n_reads, max_bytes=int(hb.seq.shape[0]) + 64,
                                  max_name_bytes=int(hb.names.shape[0]) + 64, n_slots=1, emit_fastq=True,
                                  want_matches=False, drop_bins=drop) as tx:
                        rt = tx.run(hb)
                    checked = 0
                    for b_ in range(0, rt.bin_offsets.shape[0] - 1, 8):
                        a0, a1 = int(rg.bin_offsets[b_]), int(rg.bin_offsets[b_ + 1])
                        t0_, t1_ = int(rt.bin_offsets[b_]), int(rt.bin_offsets[b_ + 1])
                        if t1_ == t0_:
                            assert a1 == a0
                            continue
                        dz = zlib.decompressobj(31)
                        if dz.decompress(rg.fastq[a0:a1].tobytes()) != rt.fastq[t0_:t1_].tobytes() or not dz.eof:
                            raise SystemExit("gzip member of bin %d does not inflate to the bin's text" % b_)
                        checked += 1
                    del rg, rt
                    for _ in range(3):
                        gz.launch(0)
                        gz.sync(0)
                    tg = gz.timings(0)
                text = tg["emit_bytes"] // 2
                out["gzip_stage"] = {"ms": tg["gzip_ms"], "text_bytes": text, "member_bytes": tg["gzip_bytes"],
                                     "ratio": tg["gzip_bytes"] / max(text, 1), "emit_ms": tg["emit_ms"],
                                     "text_gbs": text / (tg["gzip_ms"] * 1e-3) / 1e9 if tg["gzip_ms"] > 0 else None,
                                     "bins_inflated_and_compared": checked,
                                     "note": "gz_hist / gz_table / gz_measure / gz_layout / gz_zero / gz_encode "
                                             "(csrc/orc_gz.cuh) behind one step of %d reads, device time between "
                                             "events; shared-memory-pipe bound (one code and one CRC look-up per "
                                             "byte), hidden behind the copies in e2e" % hb.n_reads}
            out["counts"] = counts
        else:
            out["e2e"] = None
            out["counts"] = None
        return out

    main_cfg = run_config(args.config, args.reads, args.len_min, args.len_max, args.steps, args.warmup, True)
    t = main_cfg["timings"]
    rounds = main_cfg["rounds"]

    # ---- the only collective: per-bin count gather
    counts_np = main_cfg["counts"]
    counts = torch.from_numpy(counts_np if counts_np is not None else np.zeros(1, np.int64)).cuda()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_reads_binned = int(counts.sum().item())
    checks = None
    if main_cfg["oracle_check"] is not None:
        oc = main_cfg["oracle_check"]
        flags = torch.tensor([int(oc["bins_equal"] and oc["counts_equal"]), oc["reads"]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.SUM)
        checks = {"ranks_equal_to_oracle": int(flags[0].item()), "ranks": world, "reads_checked": int(flags[1].item()),
                  "rank0": oc}

    # ---- host link ceiling under the same N ranks
    link = measure_link(torch, barrier) if not args.no_e2e else None
    if link is not None and main_cfg["e2e"] is not None:
        e = main_cfg["e2e"]
        per_rank_s = e["ms_per_step"] * 1e-3
        e["link"] = dict(link, h2d_used_gbs=e["h2d_bytes_per_step"] / per_rank_s / 1e9,
                         d2h_used_gbs=e["d2h_bytes_per_step"] / per_rank_s / 1e9,
                         frac_of_both=max(e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]) / per_rank_s / 1e9 / link["both_gbs"],
                         note="ceilings = this rank's pinned 512 MiB copies while all %d ranks copy at once (GB/s per "
                              "direction); frac_of_both = the busier direction of the e2e leg / the bidirectional ceiling" % world)

    roofline = None
    extra = {}
    if rank == 0:
        alu_peak, sm_clk = E.measure_int32_peak(local_rank, 0)
        mix_peak, _ = E.measure_int32_peak(local_rank, 1)
        hbm, how = hbm_peak()
        m_rows = [float(np.mean([len(s) for s in r.sequences])) for r in rounds]
        # seed probes: one per base of the reads entering the round (both directions in one pass)
        seed_cols = [t["cells"][r] / (2.0 * sum(len(s) for s in rounds[r].sequences)) if rounds[r].sequences else 0.0
                     for r in range(len(rounds))]
        table = kernel_table(t, rounds, alu_peak, hbm, how, m_rows, seed_cols)
        extra["roofline_kernels"] = table
        rated = [r for r in table if r["frac"] is not None]
        top = max(rated, key=lambda r: r["ms"]) if rated else None
        cells = float(sum(t["cells"]))
        executed = float(sum(t["cells_executed"]))
        dp_ms = float(sum(t["scan_ms"]) + sum(t["trigger_ms"]))
        if top is not None:
            roofline = {"bound": top["bound"], "kernel": top["kernel"], "achieved": top["achieved"], "peak": top["peak"],
                        "unit": top["unit"], "frac": top["frac"], "ms": top["ms"], "share_of_step": top["share"],
                        "traffic": None,
                        "traffic_how": "no ncu capture on file for this kernel at this batch size (profiles/r2y_traffic.json)",
                        "how": top.get("how"),
                        "peak_how": "LOP3 issue rate %.4g lane-op/s measured in this run on this GPU (orc_measure_int32_peak "
                                    "mode 0; LOP3+IMAD mix %.4g); HBM: %s" % (alu_peak, mix_peak, how),
                        "executed": {"cells_per_step": executed, "gcups": executed / (dp_ms * 1e-3) / 1e9 if dp_ms else None,
                                     "note": "DP cells stages 1-2b really update per step, over their summed time"},
                        "algorithmic_speedup": cells / executed if executed else None,
                        "algorithmic_cells_per_step": cells}
            # DRAM bytes per launch from the committed ncu capture of the same kernel at the same batch size
            try:
                with open(os.path.join(ROOT, "profiles", "r2y_traffic.json")) as fh:
                    cap = json.load(fh)
                k = cap["kernels"].get(top["kernel"])
                if k and cap["reads_per_launch"] == args.reads and args.config in (2, 5):
                    roofline["traffic"] = k["dram_read"] + k["dram_write"]
                    roofline["traffic_how"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch, %s; an INT32-bound "
                                               "kernel: %.0f GB/s of DRAM traffic while it runs" %
                                               (cap["how"], roofline["traffic"] / (k["duration_us"] * 1e-6) / 1e9))
            except (OSError, ValueError, KeyError):
                pass
        extra["gcups"] = cells * args.steps * world / (main_cfg["ms_per_step"] * args.steps * 1e-3) / 1e9

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config in (1, 2, 3):
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core
        seed = {1: 1001, 2: 1002, 3: 1003}[args.config]
        n_s = args.cpu_sample or (16384 if args.config != 3 else 4096)
        rs_cpu = synth.generate(n_s, args.len_min, args.len_max, seed=seed, workers=min(8, ncpu))
        sub_n, times, kind, how_cpu = cpu_arm(rs_cpu, n_s, ncpu, 2, 1, 1 if args.config == 1 else 2)
        cpu_baseline = {"value": sub_n / float(np.mean(times)), "unit": UNIT, "cores": ncpu, "kind": kind,
                        "sample": "first %d reads of the workload, one warm-up and two timed passes; %s" % (sub_n, how_cpu)}

    # ---- short runs of the other single-GPU configs, so that they are driver-run numbers too
    if world == 1 and args.config == 2 and not args.no_extra:
        ex = {}
        for cfg, reads, lo, hi in ((3, 1 << 18, 1000, 3500), (4, 1 << 20, 300, 900)):
            r = run_config(cfg, reads, lo, hi, 5, 3, False)
            ex["configs[%d]" % (cfg - 1)] = {
                "workload": {3: "two-round demux on %d synthetic rRNA-cistron reads (%d-%d nt), seed 1003",
                             4: "anchored --no-indels Hamming path, 24 M13 variable indices, %d reads (%d-%d nt), seed 1004"}[cfg]
                            % (reads, lo, hi),
                "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": 5, "warmup": 3,
                "e2e": r["e2e"], "gpu_launches": r["launches"],
                "stages_ms": {k: r["timings"][k] for k in ("pack_ms", "trigger_ms", "scan_ms", "resolve_ms", "bin_ms", "emit_ms", "total_ms")}}
        extra["extra_configs"] = ex

    if rank == 0:
        reads = args.reads
        stage = {"pack_ms": t["pack_ms"], "trigger_ms": t["trigger_ms"], "scan_ms": t["scan_ms"], "resolve_ms": t["resolve_ms"],
                 "bin_ms": t["bin_ms"], "emit_ms": t["emit_ms"], "total_ms": t["total_ms"], "n_tasks": t["n_tasks"],
                 "n_tasks_wide": t["n_tasks_wide"], "n_pairs_2b": t["n_pairs_2b"], "kernel_ms": t["kernel_ms"]}
        workload = {1: "configs[0]: round-1 SP5 5' demux + trim only, %d synthetic reads (%d-%d nt) per GPU, "
                       "-g file:M13_amplicon_indices_forward.fa -e 0.1 --rc",
                    2: "configs[1]: full two-round SP5->SP27 combinatorial demux + trim on %d synthetic "
                       "COI-length reads (%d-%d nt) per GPU, -e 0.1 -O 3 --rc, 12+12 M13 indices",
                    3: "configs[2]: two-round demux on %d synthetic rRNA-cistron reads (%d-%d nt) per GPU",
                    4: "configs[3]: anchored --no-indels Hamming path, 24 M13 variable indices, %d reads "
                       "(%d-%d nt) per GPU, index at read offset 0",
                    5: "configs[4]: two-round demux sharded over the GPUs, %d reads (%d-%d nt) per shard and step, "
                       "distinct shards made on the GPU"}[args.config] % (reads, args.len_min, args.len_max)
        cfg = {"workload": workload, "reads_per_gpu_per_step": reads, "seed": main_cfg["seed"],
               "l2": "inputs larger than L2 (%.2f GB of inputs and packed codes resident per step)"
                     % (2.5 * (main_cfg["n_bytes"] or reads * (args.len_min + args.len_max) // 2) / 1e9),
               "parallelism": "reads sharded by batch, no data-path collective; all_reduce of %d bin counters at the end"
                              % int(counts.numel())}
        cfg["resident_batches_in_flight"] = ("%d batches resident per GPU, step k runs on slot k %% %d; every slot has its own "
                                             "stream, so consecutive steps overlap (as in the submit/wait pipeline); timed "
                                             "with events that bracket all streams" % (main_cfg["n_slots"], main_cfg["n_slots"]))
        if args.config == 5:
            distinct = main_cfg["n_slots"] * world * reads
            cfg["distinct_reads_total"] = distinct
            cfg["reads_processed_in_timed_region"] = args.steps * world * reads
            cfg["target_reads"] = TARGET_READS
        line = {
            "metric": METRIC_R1 if args.config == 1 else METRIC, "value": main_cfg["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_cfg["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": cfg, "clocks": main_cfg["clocks"], "e2e": main_cfg["e2e"], "gpu_launches": main_cfg["launches"],
            "roofline": roofline, "cpu_baseline": cpu_baseline, "stages_ms_last_step": stage,
            "wall_ms_per_step": main_cfg["wall_ms_per_step"], "reads_binned_all_ranks": total_reads_binned,
            "oracle_check": checks, "gen_s": main_cfg["gen_s"], "cpu_affinity": numa,
            "verify_open": "parity unpinned: no real cutadapt 4.9 here; SURVEY VERIFY-1..15 stay open until "
                           "tests/test_cutadapt_diff.py runs somewhere (it skips without cutadapt)",
        }
        if main_cfg.get("e2e_text") is not None:
            line["e2e_text"] = main_cfg["e2e_text"]
        if main_cfg.get("gzip_stage") is not None:
            line["gzip_stage"] = main_cfg["gzip_stage"]
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
