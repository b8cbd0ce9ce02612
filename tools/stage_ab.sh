#!/bin/bash
# A/B of the staged-copy policy on ONE box (host performance differs a lot between boxes): usage: bash tools/stage_ab.sh
for cfg in "SPX_STAGE_LANES=0" "SPX_STAGE_LANES=4" "SPX_STAGE_LANES=6" "SPX_STAGE_LANES=8" "SPX_STAGE_LANES=3"; do
  echo "== $cfg"
  env $cfg python tools/click_stages.py 3840 2160 | awk '{printf "%s %s | ", $(NF-2)=="total"?"total":$1, $(NF-1)} END {print ""}'
  env $cfg python tools/click_stages.py | awk '{printf "%s %s | ", $(NF-2)=="total"?"total":$1, $(NF-1)} END {print ""}'
done
nproc
