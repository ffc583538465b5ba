"""Summarise ncu outputs for profiles/: launch list shares and key metrics of a .ncu-rep."""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for k, v in agg.items() if 'int32_peak' not in k and 'at::' not in k)
    print("| kernel | launches | total ms | share of step kernels |\n|---|---|---|---|")
    for k, v in agg.items():
        if 'int32_peak' in k or 'at::' in k:
            continue
        print("| %s | %d | %.3f | %.1f%% |" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("\n### %s" % r[h.index("Kernel Name")].split('(')[0])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print("- %s = %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        rep(sys.argv[2])
