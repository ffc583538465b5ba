#!/usr/bin/env python
"""All-vs-all edit distances of synthetic reads on one GPU (amplicon_sorter's
compare-all step, amplicon_sorter.py:225-235), beside the oracle's scalar recurrence on the host
cores (NOT edlib: a plain O(nm) DP, so the CPU figure is a lower bound of what edlib would do).

    python tools/edit_bench.py [--reads 2000] [--len-min 300] [--len-max 900] [--mode NW]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000)
    ap.add_argument("--len-min", type=int, default=300)
    ap.add_argument("--len-max", type=int, default=900)
    ap.add_argument("--mode", default="NW", choices=["NW", "HW"])
    ap.add_argument("--cpu-pairs", type=int, default=20000)
    a = ap.parse_args()
    import numpy as np
    import oracle
    from orcdemux import distance as D
    from orcdemux import synth
    rs = synth.generate(a.reads, a.len_min, a.len_max, seed=1003, workers=min(8, os.cpu_count() or 1))
    pa, pb = D.all_pairs(rs.n_reads)
    L = rs.lengths.astype(np.float64)
    cells = float((np.minimum(L[pa], L[pb]) * np.maximum(L[pa], L[pb])).sum())
    D.edit_distances(rs.seq, rs.offsets, rs.lengths, pa[:1000], pb[:1000], a.mode)         # warm-up
    t0 = time.perf_counter()
    dist, ms = D.edit_distances(rs.seq, rs.offsets, rs.lengths, pa, pb, a.mode, with_time=True)
    wall = time.perf_counter() - t0
    k = min(a.cpu_pairs, pa.shape[0])
    sel = np.random.default_rng(1).choice(pa.shape[0], k, replace=False)
    ncpu = os.cpu_count() or 1
    t0 = time.perf_counter()
    exp = oracle.edit_distances(rs.seq, rs.offsets, rs.lengths, pa[sel], pb[sel], a.mode, n_threads=ncpu)
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(exp, dist[sel]), "kernel and recurrence differ"
    cpu_cells = float((np.minimum(L[pa[sel]], L[pb[sel]]) * np.maximum(L[pa[sel]], L[pb[sel]])).sum())
    from orcdemux import engine as E
    alu_peak, _ = E.measure_int32_peak(0, 0)
    # 27 ALU-pipe instructions per 64-row block and column in edit_kernel's loop (SASS: 17 LOP3, 3 IADD3,
    # 2 SHF, 2 SEL, 2 ISETP, 1 LEA) + about 3 of the step's own -> 30
    peak_gcups = alu_peak / 30.0 * 64.0 / 1e9
    print(json.dumps({
        "what": "all-vs-all %s edit distance, %d synthetic reads %d-%d nt" % (a.mode, rs.n_reads, a.len_min, a.len_max),
        "pairs": int(pa.shape[0]), "kernel_ms": ms, "pairs_per_s": pa.shape[0] / (ms * 1e-3),
        "gcups": cells / (ms * 1e-3) / 1e9, "call_wall_s": wall,
        "roofline": {"bound": "int32_alu", "achieved": cells / (ms * 1e-3) / 1e9, "peak": peak_gcups, "unit": "GCUPS",
                     "frac": cells / (ms * 1e-3) / 1e9 / peak_gcups,
                     "peak_how": "measured LOP3 rate %.3g lane-op/s / 30 ALU-pipe instr per block step x 64 cells, "
                                 "every lane busy; the kernel leaves lanes idle while a pair's wavefront fills and "
                                 "drains and where a query has fewer blocks than its 8/16/32 lanes" % alu_peak},
        "cpu_port": {"pairs": int(k), "seconds": cpu_s, "pairs_per_s": k / cpu_s, "gcups": cpu_cells / cpu_s / 1e9,
                     "cores": ncpu, "note": "oracle/edit_oracle.c scalar recurrence, not edlib"},
        "checked_against_oracle": int(k)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
