#!/usr/bin/env python
"""Per-sub-batch device times inside the pipelined e2e leg (orc_submit / orc_wait over S slots), with the
qualities copied or read in place: where does a step's time go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200")]
import numpy as np
import bench
from orcdemux import engine as E, synth

COMBOS = [tuple(int(v) for v in c.split("x")) for c in os.environ.get("COMBOS", "4x8").split(",")]
STEPS = 6
rs = E.pin_readset(synth.generate(1 << 20, 300, 900, seed=1002, workers=16))
drop = bench.scripts_final_tree_drop()
if os.environ.get('DROPALL'):
    drop[:] = 1             # nothing emitted: the H2D side and the matching kernels alone
GZ = bool(os.environ.get('GZ'))
for S, NSUB, zc in [(s_, n_, z) for s_, n_ in COMBOS for z in ((False, True) if os.environ.get('BOTH') else (True,))]:
    subs, per = bench.split_subbatches(E, synth, rs, NSUB)
    eng = E.Engine(E.m13_rounds(), device=0, max_reads=per, max_bytes=max(int(x.seq.shape[0]) for x in subs) + 64,
                   max_name_bytes=max(int(x.names.shape[0]) for x in subs) + 64, n_slots=S, emit_fastq=True,
                   want_matches=False, drop_bins=drop, qual_zero_copy=zc, emit_gzip=GZ)
    acc = {"h2d_ms": [], "total_ms": [], "emit_ms": [], "d2h_ms": [], "host_submit_ms": [], "host_wait_ms": []}
    tl = []
    infl = []
    full_tl = []
    w_ref = time.perf_counter()
    def take(slot):
        t0 = time.perf_counter(); eng.wait(slot, copy=False); t1 = time.perf_counter(); acc["host_wait_ms"].append(1e3 * (t1 - t0))
        if os.environ.get('FULLTL'):
            full_tl.append(eng.timeline(slot) + [1e3 * (t0 - w_ref), 1e3 * (t1 - w_ref)])
        # (no eng.timings() here: its synchronous cudaMemcpy of the counters queues behind the pending copies
        # of the other slots and stalls this thread -- the timeline below is read once, after the loop)
    k = 0
    if os.environ.get('THREADS'):
        # one thread submits, another waits: ctypes drops the GIL inside the calls, so the D2H side no longer holds
        # up the next upload (slots are handed over through two queues)
        import queue, threading
        free, busy = queue.Queue(), queue.Queue()
        for sl in range(S):
            free.put(sl)
        def waiter():
            while True:
                sl = busy.get()
                if sl is None:
                    return
                eng.wait(sl, copy=False)
                free.put(sl)
        for rep in range(STEPS + 1):
            if rep == 1:
                busy.join() if False else None
                while free.qsize() < S:
                    time.sleep(0.0005)
                import torch; torch.cuda.synchronize(); w0 = time.perf_counter()
            if rep == 0:
                th = threading.Thread(target=waiter); th.start()
            for sub in subs:
                sl = free.get()
                eng.submit(sl, sub)
                busy.put(sl)
        while free.qsize() < S:
            time.sleep(0.0005)
        secs = time.perf_counter() - w0
        busy.put(None); th.join()
        print("S=%d NSUB=%d zero_copy=%s THREADS  %.2f ms/step  %.1f M reads/s" % (S, NSUB, zc, 1e3 * secs / STEPS, STEPS * rs.n_reads / secs / 1e6))
        eng.close()
        continue
    for rep in range(STEPS + 1):
        if rep == 1:
            while infl: take(infl.pop(0))
            for v in acc.values(): v.clear()
            import torch; torch.cuda.synchronize(); w0 = time.perf_counter()
        for sub in subs:
            if len(infl) == S: take(infl.pop(0))
            t0 = time.perf_counter(); eng.submit(k % S, sub); acc["host_submit_ms"].append(1e3 * (time.perf_counter() - t0))
            infl.append(k % S); k += 1
    while infl: take(infl.pop(0))
    secs = time.perf_counter() - w0
    print("S=%d NSUB=%d " % (S, NSUB) + "zero_copy=%s  %.2f ms/step  %.1f M reads/s  per sub-batch (mean ms): %s" % (
        zc, 1e3 * secs / STEPS, STEPS * rs.n_reads / secs / 1e6,
        "  ".join("%s %.3f" % (n, float(np.mean(v))) for n, v in acc.items())))
    for slot in range(S):
        t = eng.timings(slot)
        for k_ in ("h2d_ms", "total_ms", "emit_ms", "d2h_ms"): acc[k_].append(t[k_])
        tl.append((slot, t["timeline_ms"]))
    tl.sort(key=lambda e: e[1][0])
    if os.environ.get('TIMELINE'):
        t0 = tl[0][1][0]
        for slot, x in tl:
            print('   slot %d  h2d %7.2f-%7.2f  kernels -%7.2f  emit -%7.2f  d2h -%7.2f' % (slot, x[0] - t0, x[1] - t0, x[2] - t0, x[3] - t0, x[4] - t0))
    if os.environ.get('FULLTL'):
        # every sub-batch of the last steps in the order they were waited for: when its stages ended on the device
        # (ms since the first one shown), how long each engine was busy between consecutive ends, and when the host
        # entered / left orc_wait (host clock, own origin)
        rows = full_tl[-3 * NSUB:]
        o = rows[0][0]
        print('   sub   h2d_start  h2d_end   match_end  emit_end   gz_end    d2h_end  | h2d_end-prev  d2h_end-prev | host wait in..out')
        for i, x in enumerate(rows):
            p_ = rows[i - 1] if i else x
            print('   %3d  %9.2f %9.2f %9.2f %9.2f %9.2f %9.2f  | %8.2f %12.2f   | %9.2f %9.2f' % (
                i, x[0] - o, x[1] - o, x[2] - o, x[3] - o, x[4] - o, x[5] - o, x[1] - p_[1], x[5] - p_[5], x[6], x[7]))
    eng.close()
