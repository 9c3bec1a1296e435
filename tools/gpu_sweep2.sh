#!/bin/bash
# Sweep blend_bwd variants (pixels/thread x min CTAs per SM) at C2.
for cfg in "4 0" "4 10" "4 12" "2 0" "2 6" "2 8"; do
  set -- $cfg
  export GSR_BWD_PPT=$1 GSR_BWD_MINB=$2
  echo -n "BWD_PPT=$1 MINB=$2: "
  timeout 200 python bench.py --views 2 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py | grep -E "blend_bwd" 
done
