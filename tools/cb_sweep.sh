for cfg in "8 16 512" "8 16 1024" "8 24 1024" "8 12 512"; do set -- $cfg
SPX_CB_WARM=$1 SPX_CB_ROWS=$2 SPX_CB_T=$3 SPX_CONTOUR_DEBUG=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv python tools/prof_post.py 2>&1 | grep contour | tr '\n' ' '
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
print("warm $1 rows $2 T $3:", [float(r[iv])/1000 for r in rows[1:] if "contour" in r[ik]])
PY
done
