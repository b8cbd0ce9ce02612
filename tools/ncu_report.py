#!/usr/bin/env python
"""Summarise an .ncu-rep (raw metrics + per-phase hot spots) -- used to write profiles/*.txt.
usage: python tools/ncu_report.py gpurun_out/prof.ncu-rep [N_top_groups]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 10
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')][:100])
    for k in keys:
        if k in hdr:
            print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and 'per_warp_active' in h and 'not_issued' not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 3:
                print("  stall %-64s %.1f %%" % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
body = [r for r in rows[idx[0] + 2: idx[1] if len(idx) > 1 else None] if len(r) > 5 and r[5].isdigit()]
tot_i = sum(int(r[5]) for r in body); tot_s = sum(int(r[4]) for r in body)
print("source page: %d SASS instructions, %d executed warp-instructions, %d samples" % (len(body), tot_i, tot_s))
groups = []
for r in body:
    c = int(r[5])
    if groups and abs(groups[-1][0] - c) <= max(2, 0.001 * c): groups[-1][1].append(r)
    else: groups.append([c, [r]])
for c, g in sorted(groups, key=lambda g: -sum(int(r[4]) for r in g[1]))[:ntop]:
    print("== block @%s..%s  %d instr x %d executions = %.1f%% of instructions, %.1f%% of samples" % (
        g[0][0][-5:], g[-1][0][-5:], len(g), c, 100.0 * c * len(g) / tot_i, 100.0 * sum(int(r[4]) for r in g) / tot_s))
    top = sorted(g, key=lambda r: -int(r[4]))[:4]
    for r in top:
        print("      %-70s samples %s" % (r[1].strip()[:70], r[4]))
