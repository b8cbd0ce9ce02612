#!/bin/bash
# per-launch device times of a command, caches left warm between launches (ncu --cache-control none): usage: launch_times.sh <filter-regex> <cmd...>
f=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/lt.csv "$@" > /dev/null 2>&1
python - "$f" <<'PY'
import csv, re, sys
rows=[r for r in csv.reader(open("gpurun_out/lt.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]:
    if re.search(sys.argv[1], r[ik]): print("%-70s %8.2f us" % (r[ik][:70], float(r[iv])/1000))
PY
