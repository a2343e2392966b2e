#!/usr/bin/env python
"""Summarise ncu exports kept under profiles/.

  summarize.py launches <launches.csv>          per-kernel totals / shares of an
                                                `ncu --metrics gpu__time_duration.sum --csv` log
  summarize.py raw <ncu --page raw --csv file>  key counters per profiled launch, as markdown
"""
import collections
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'total ms':>10s} {'avg ms':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:72s} {len(v):4d} {sum(v)/1e6:10.3f} {sum(v)/len(v)/1e6:10.3f} {100*sum(v)/tot:6.1f}%")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"### {d['Kernel Name']}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {u[k]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
