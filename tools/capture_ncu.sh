#!/bin/bash
# ncu --set full captures of the path's kernels (one launch each), reports into gpurun_out/.
# usage (on the GPU box): bash tools/capture_ncu.sh [slic|post|lsc|seeds|cvt ...]
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for what in "$@"; do
case $what in
slic)  python tools/prof_slic.py > gpurun_out/plain_slic.log 2>&1 && $NCU -k regex:slic_tile_kernel -s 5 -c 1 -f -o gpurun_out/r2_slic_tile python tools/prof_slic.py > gpurun_out/ncu_slic.log 2>&1
       $NCU -k regex:'slic_finalize|slic_seed|slic_expand|slic_bin' -s 0 -c 5 -f -o gpurun_out/r2_slic_small python tools/prof_slic.py >> gpurun_out/ncu_slic.log 2>&1 ;;
post)  python tools/prof_post.py > gpurun_out/plain_post.log 2>&1 && $NCU -k regex:'ccl_|scan|contour_' -c 24 -f -o gpurun_out/r2_post python tools/prof_post.py > gpurun_out/ncu_post.log 2>&1 ;;
cvt)   $NCU -k regex:cvt8 -c 1 -f -o gpurun_out/r2_cvt8 python tools/prof_post.py > gpurun_out/ncu_cvt.log 2>&1 ;;
lsc)   python tools/prof_lsc.py 30 3840 2160 > gpurun_out/plain_lsc.log 2>&1 && $NCU -k regex:'lsc_tile_kernel|lsc_fixpoint|lsc_rounds' -s 3 -c 2 -f -o gpurun_out/r2_lsc python tools/prof_lsc.py 30 3840 2160 > gpurun_out/ncu_lsc.log 2>&1 ;;
seeds) python tools/prof_seeds.py > gpurun_out/plain_seeds.log 2>&1 && SPX_NO_GRAPH=1 $NCU -k regex:'seeds_(block|pixel)' -s 60 -c 8 -f -o gpurun_out/r2_seeds python tools/prof_seeds.py > gpurun_out/ncu_seeds.log 2>&1 ;;
lscpost) $NCU -k regex:'lsc_fixpoint|lsc_rounds|lsc_edge|lsc_post' -c 8 -f -o gpurun_out/r2_lsc_post python tools/prof_lsc.py 30 3840 2160 > gpurun_out/ncu_lscpost.log 2>&1 ;;
esac
done
ls -la gpurun_out/*.ncu-rep
