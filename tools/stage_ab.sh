for cfg in "SPX_STAGE_LANES=0" "SPX_STAGE_MIN_CHUNK_KB=2048 SPX_STAGE_PER_LANE_MB=2" "SPX_STAGE_MIN_CHUNK_KB=256 SPX_STAGE_PER_LANE_MB=4" "SPX_STAGE_MIN_CHUNK_KB=512 SPX_STAGE_PER_LANE_MB=2" "SPX_STAGE_LANES=2 SPX_STAGE_MIN_CHUNK_KB=2048 SPX_STAGE_PER_LANE_MB=2" "SPX_STAGE_LANES=8 SPX_STAGE_MIN_CHUNK_KB=2048 SPX_STAGE_PER_LANE_MB=2"; do
  echo "== $cfg"
  env $cfg python tools/click_stages.py 3840 2160 | awk '{printf "%s %s | ", $(NF-2)=="total"?"total":$1, $(NF-1)} END {print ""}'
  env $cfg python tools/click_stages.py | awk '{printf "%s %s | ", $(NF-2)=="total"?"total":$1, $(NF-1)} END {print ""}'
done
