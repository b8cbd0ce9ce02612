#!/usr/bin/env python
"""One line per profiled launch of an .ncu-rep (ncu --set full): duration, DRAM bytes and achieved GB/s, executed warp
instructions, issue utilisation, achieved occupancy, launch geometry.  usage: python tools/ncu_brief.py file.ncu-rep [...]"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread']
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    print("== %s   (times under ncu: cold caches, serialised launches)" % rep)
    for r in rows[2:]:
        d = {k: r[h.index(k)] for k in KEYS if k in h}
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tb = sum(float(d[k]) * scale[u[h.index(k)]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        tus = float(d['gpu__time_duration.sum']) * {'us': 1, 'ms': 1e3, 'ns': 1e-3, 's': 1e6}[u[h.index('gpu__time_duration.sum')]]
        print('%-58s %8.1f us  dram %7.2f MB = %5.0f GB/s  warp-inst %10.0f  issue %5.1f %%  warps active %5.1f %%  grid %s x %s  regs %s' % (
            r[h.index('Kernel Name')][:58], tus, tb / 1e6, tb / tus / 1e3, float(d['smsp__inst_executed.sum']),
            float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']), float(d['sm__warps_active.avg.pct_of_peak_sustained_active']),
            d['launch__grid_size'], d['launch__block_size'], d['launch__registers_per_thread']))
