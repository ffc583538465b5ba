#!/usr/bin/env python
"""bench.py -- reads/s of the two-round SP5 x SP27 demultiplex (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (B200, liborcdemux.so)
    python bench.py --impl reference [...]                       the reference's CPU path

A "step" is one pass of the whole hot path (pack, round-1 scan/resolve/select, round-2
scan/resolve/select, bin partition, FASTQ emit) over one batch of synthetic reads.  At N=1
the workload is BASELINE configs[1]: 1 Mi synthetic COI-length reads (300-900 nt, seed 1002),
full two-round demux + trim.  With N>1 (torchrun, one rank per GPU) every rank runs its own
shard of the same size (weak scaling, seeds (1005<<32)+rank as in SURVEY 8d config 5); the
only collective is the final per-bin count gather (all_reduce of n_bins counters).

`value`   reads/s with the batch resident in HBM (kernels only), CUDA events on the library's
          stream, max over ranks.
`e2e`     the same metric through the C ABI with host buffers: every step copies the inputs
          from pinned host memory and the results (FASTQ text, bins, matches) back.
`roofline` for the dominant kernel (the bit-parallel scan): algorithmic DP cells per second
          against the INT32 ALU-pipe peak measured on this GPU by orc_measure_int32_peak().
`cpu_baseline` / `--impl reference`: the reference's own implementation of this path is
          cutadapt 4.9, which is neither vendored in the reference nor installable here, so the
          CPU arm is the oracle's C restatement of it ("port"), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
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
# ALU-pipe instructions per DP column of one (read, adapter, orientation) pair in
# scan_kernel's inner loop, counted in the SASS (profiles/README.md); a column is m cells.
SCAN_ALU_INSTR_PER_COLUMN = 24.0
# DRAM traffic of the scan stages (seed_kernel + trigger_kernel + filter_kernel + scan_kernel, both rounds)
# per read, from the ncu --set full capture profiles/r1n_main_raw.csv (dram__bytes_read.sum +
# dram__bytes_write.sum of the eight launches at 262 144 COI reads, divided by the reads)
SCAN_DRAM_BYTES_PER_READ = 3.157e3


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
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[]: 1 = round 1 only (SP5 5' demux) on 100 k reads, the reference's "
                         "own CPU-runnable case; 2 = 1 Mi COI reads two-round (default, the metric's config), "
                         "3 = rRNA-cistron reads 1-3.5 kb two-round, 4 = anchored --no-indels Hamming path "
                         "(24 M13 variable indices) on reads with the bare index at offset 0")
    ap.add_argument("--sub-batches", type=int, default=8)
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
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_arm(rs, n_sample, threads, steps, warmup, n_rounds=2):
    """Time the oracle's C restatement of cutadapt on the first n_sample reads."""
    import oracle
    from orcdemux import m13
    sub_n = min(n_sample, rs.n_reads)
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
    return sub_n, times


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

    from orcdemux import synth

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        import oracle
        oracle.build()
        threads = ncpu
        one_round = args.config == 1
        n_sample = args.cpu_sample or (100000 if one_round else 16384)
        rs = synth.generate(n_sample, args.len_min, args.len_max, seed=1001 if one_round else 1002, workers=min(8, ncpu))
        warm = min(args.warmup, 1)
        sub_n, times = cpu_arm(rs, n_sample, threads, args.steps, warm, 1 if one_round else 2)
        ms = 1e3 * float(np.mean(times))
        val = sub_n / float(np.mean(times))
        line = {
            "impl": "reference", "metric": METRIC_R1 if one_round else METRIC, "value": val, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": ("configs[0]: round-1 SP5 5' demux (-g file:M13_amplicon_indices_forward.fa -e 0.1 --rc), "
                                    "%d synthetic reads %d-%d nt, seed 1001" % (sub_n, args.len_min, args.len_max))
                       if one_round else
                       "configs[1]: two-round SP5->SP27 demux + trim, synthetic COI reads %d-%d nt, "
                       "seed 1002; each step = the first %d reads of the 1 Mi-read workload"
                       % (args.len_min, args.len_max, sub_n),
                       "reads_per_step": sub_n, "note": "cutadapt 4.9 is not vendored in the reference and not "
                       "installable here: this arm is the restated-cutadapt CPU baseline (oracle/cutadapt_oracle.c, "
                       "Ukkonen-banded DP, pthreads), not upstream cutadapt"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d reads of the workload, %d timed passes" % (sub_n, args.steps)},
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

    seed = {1: 1001, 2: 1002, 3: 1003, 4: 1004}[args.config] if world == 1 else (1005 << 32) + rank
    workers = max(1, min(16, ncpu // max(world, 1)))
    t0 = time.perf_counter()
    rs = synth.generate(args.reads, args.len_min, args.len_max, seed=seed, workers=workers,
                        anchored=(args.config == 4))
    if args.config == 4:
        from orcdemux import m13
        from orcdemux.lib import ORC_PREFIX
        var = m13.variable_all()
        rounds = [E.Round([n for n, _ in var], [q for _, q in var], ORC_PREFIX, 0.1, 3, False, True)]
    elif args.config == 1:
        rounds = E.m13_rounds()[:1]
    else:
        rounds = E.m13_rounds()
    gen_s = time.perf_counter() - t0
    rs = E.pin_readset(rs)
    n_bytes = int(rs.seq.shape[0])
    n_slots = 1
    eng = E.Engine(rounds, device=local_rank, max_reads=rs.n_reads, max_bytes=n_bytes,
                   max_name_bytes=int(rs.names.shape[0]) + 64, n_slots=n_slots, emit_fastq=True, want_matches=True)

    # ---- device-resident: `value`
    eng.upload(0, rs)
    eng.sync(0)
    for _ in range(args.warmup):
        eng.launch(0)
    eng.sync(0)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    eng.timer_start(0)
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.launch(0)
    dev_ms = eng.timer_stop(0)          # device time of exactly K steps on the library's stream
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop()
    t = eng.timings(0)                  # stage split of the last step
    cells = float(sum(t["cells"]))
    scan_ms = float(sum(t["scan_ms"]) + sum(t["trigger_ms"]))
    stage = {"pack_ms": t["pack_ms"], "trigger_ms": t["trigger_ms"], "scan_ms": t["scan_ms"], "resolve_ms": t["resolve_ms"],
             "bin_ms": t["bin_ms"], "emit_ms": t["emit_ms"], "total_ms": t["total_ms"], "n_tasks": t["n_tasks"]}
    launches = int(t["kernel_launches"]) * args.steps

    tm = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    max_ms = float(tm.item())
    value = (args.reads * world * args.steps) / (max_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers: `e2e`
    # The step's reads are streamed as SUB sub-batches through S slots (what a FASTQ reader does):
    # the H2D copy of one sub-batch overlaps the kernels of the previous and the D2H of the one
    # before.  Every input byte is copied from pinned host memory and every result byte (FASTQ
    # text, bins, lengths, match records) is copied back, every step.
    e2e = None
    eng.close()
    if not args.no_e2e:
        SUB, S = max(1, args.sub_batches), 4
        per = (args.reads + SUB - 1) // SUB
        subs = []
        for i in range(SUB):
            lo, hi = i * per, min(args.reads, (i + 1) * per)
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
        eng2 = E.Engine(rounds, device=local_rank, max_reads=per,
                        max_bytes=max(int(x.seq.shape[0]) for x in subs) + 64,
                        max_name_bytes=max(int(x.names.shape[0]) for x in subs) + 64, n_slots=S,
                        emit_fastq=True, want_matches=True)
        h2d = sum(int(x.seq.nbytes + x.qual.nbytes + x.offsets.nbytes + x.lengths.nbytes + x.names.nbytes +
                      x.name_offsets.nbytes) for x in subs)
        state = {"inflight": [], "k": 0, "reads": 0, "d2h": 0}

        def pump(sub):
            if len(state["inflight"]) == S:
                r = eng2.wait(state["inflight"].pop(0), copy=False)
                state["reads"] += int(r.bin_counts.sum())
                state["d2h"] += int(r.fastq.nbytes + r.bin.nbytes + r.out_len.nbytes + r.bin_counts.nbytes +
                                    r.bin_offsets.nbytes + sum(m.nbytes for m in r.matches))
            slot = state["k"] % S
            eng2.submit(slot, sub)
            state["inflight"].append(slot)
            state["k"] += 1

        def drain():
            while state["inflight"]:
                r = eng2.wait(state["inflight"].pop(0), copy=False)
                state["reads"] += int(r.bin_counts.sum())
                state["d2h"] += int(r.fastq.nbytes + r.bin.nbytes + r.out_len.nbytes + r.bin_counts.nbytes +
                                    r.bin_offsets.nbytes + sum(m.nbytes for m in r.matches))

        for sub in subs:                            # warm the copy paths: one untimed step
            pump(sub)
        drain()
        state.update(reads=0, d2h=0)
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            for sub in subs:
                pump(sub)
        drain()
        barrier()
        e2e_s = time.perf_counter() - w0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert state["reads"] == args.steps * args.reads
        e2e = {"value": (args.reads * world * args.steps) / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": state["d2h"] // args.steps,
               "ms_per_step": 1e3 * float(te.item()) / args.steps,
               "note": "each step streamed as %d sub-batches over %d slots/streams (copies overlap kernels); "
                       "pipeline fill and drain are inside the timed region; host clock between device syncs "
                       "because the region includes host-side calls" % (len(subs), S)}
        counts_np = eng2.counts().astype(np.int64)
        eng2.close()
    else:
        counts_np = None

    # ---- the only collective: per-bin count gather
    counts = torch.from_numpy(counts_np if counts_np is not None else np.zeros(1, np.int64)).cuda()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_reads_binned = int(counts.sum().item())

    # ---- roofline of the dominant kernel
    roofline = None
    cpu_baseline = None
    extra = {}
    if rank == 0:
        alu_peak, sm_clk = E.measure_int32_peak(local_rank, 0)
        mix_peak, _ = E.measure_int32_peak(local_rank, 1)
        m_rows = 59.0
        peak_gcups = alu_peak / SCAN_ALU_INSTR_PER_COLUMN * m_rows / 1e9
        ach_gcups = cells / (scan_ms * 1e-3) / 1e9
        exe_gcups = float(sum(t["cells_executed"])) / (scan_ms * 1e-3) / 1e9
        roofline = {"bound": "int32_alu",
                    "kernel": "scan = seed_kernel + trigger_kernel (stage 1) + filter_kernel (stage 2a) + scan_kernel (stage 2b), both rounds",
                    "achieved": ach_gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": ach_gcups / peak_gcups,
                    "traffic": SCAN_DRAM_BYTES_PER_READ * args.reads if args.config == 2 else None,
                    "traffic_how": "ncu dram bytes of the scan launches per read (profiles/r1n_main_raw.csv) x reads; "
                                   "ALU-bound kernels: traffic is the packed codes read once per stage, far below HBM limits",
                    "executed": {"achieved": exe_gcups, "frac": exe_gcups / peak_gcups,
                                 "note": "DP cells the kernels really update; the rest of the algorithmic cells "
                                         "(2*12*m*n per read and round, SURVEY 8d) are skipped exactly by the "
                                         "seed filter (stage 1, no DP cells at all in its main pass) and the "
                                         "32-row block test (stage 2a), which is why `frac` exceeds 1"},
                    "peak_how": "measured LOP3 issue rate %.3g lane-op/s (orc_measure_int32_peak mode 0, this GPU, this "
                                "run) / %.0f ALU-pipe instr per 64-bit Myers column x %d rows = what an exhaustive "
                                "per-pair scan can reach; LOP3+IMAD mix: %.3g" %
                                (alu_peak, SCAN_ALU_INSTR_PER_COLUMN, int(m_rows), mix_peak),
                    "cells_per_step": cells, "scan_ms": scan_ms}
        hbm, how = hbm_peak()
        extra["roofline_hbm"] = [
            {"kernel": "pack_kernel", "bound": "hbm", "achieved": t["pack_bytes"] / (t["pack_ms"] * 1e-3) / 1e9,
             "peak": hbm, "unit": "GB/s", "frac": t["pack_bytes"] / (t["pack_ms"] * 1e-3) / 1e9 / hbm,
             "peak_how": "of " + how},
            {"kernel": "emit_kernel", "bound": "hbm", "achieved": t["emit_bytes"] / (t["emit_ms"] * 1e-3) / 1e9,
             "peak": hbm, "unit": "GB/s", "frac": t["emit_bytes"] / (t["emit_ms"] * 1e-3) / 1e9 / hbm,
             "peak_how": "of " + how}]
        extra["gcups"] = cells * args.steps * world / (max_ms * 1e-3) / 1e9 if world == 1 else None

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config != 4:
        import oracle
        oracle.build()
        n_s = args.cpu_sample or 16384
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core
        sub_n, times = cpu_arm(rs, n_s, ncpu, 2, 1, 1 if args.config == 1 else 2)
        cpu_baseline = {"value": sub_n / float(np.mean(times)), "unit": UNIT, "cores": ncpu, "kind": "port",
                        "sample": "first %d reads of the workload, one warm-up and two timed passes, %d threads "
                                  "(restated-cutadapt CPU baseline, not upstream cutadapt)" % (sub_n, ncpu)}

    if rank == 0:
        line = {
            "metric": METRIC_R1 if args.config == 1 else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": {1: "configs[0]: round-1 SP5 5' demux + trim only, %d synthetic reads (%d-%d nt) per GPU, "
                                       "-g file:M13_amplicon_indices_forward.fa -e 0.1 --rc",
                                    2: "configs[1]: full two-round SP5->SP27 combinatorial demux + trim on %d synthetic "
                                       "COI-length reads (%d-%d nt) per GPU, -e 0.1 -O 3 --rc, 12+12 M13 indices",
                                    3: "configs[2]: two-round demux on %d synthetic rRNA-cistron reads (%d-%d nt) per GPU",
                                    4: "configs[3]: anchored --no-indels Hamming path, 24 M13 variable indices, %d reads "
                                       "(%d-%d nt) per GPU, index at read offset 0"}[args.config]
                                   % (args.reads, args.len_min, args.len_max),
                       "reads_per_gpu": args.reads, "seed": seed, "l2": "inputs larger than L2 (%.2f GB resident "
                       "per step)" % (2.5 * n_bytes / 1e9), "parallelism": "reads sharded by batch, no data-path "
                       "collective; all_reduce of %d bin counters at the end" % int(counts.numel())},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "stages_ms_last_step": stage, "wall_ms_per_step": wall_ms / args.steps,
            "reads_binned_all_ranks": total_reads_binned, "gen_s": gen_s, "cpu_affinity": numa,
        }
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
