#!/bin/bash
# Sweep blend-kernel variants: correctness (vs reference) and per-kernel time at C2.
mkdir -p gpurun_out
for cfg in "1 1" "2 2" "4 4" "1 2" "2 4"; do
  set -- $cfg
  export GSR_FWD_PPT=$1 GSR_BWD_PPT=$2
  echo "=== FWD_PPT=$1 BWD_PPT=$2"
  timeout 200 python tools/gpu_diag.py --P 200000 --W 800 --H 600 2>&1 | grep -E '"name": "(color|n_contrib|final_T|grad_means3D|grad_scales|grad_shs|se3 dS)"' | grep -E "rel_to_max|bit_mismatch" | cut -c1-150
  timeout 200 python bench.py --views 2 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('ms_per_view %.3f  blend_fwd %.3f  blend_bwd %.3f  pre_bwd %.3f sort_pass %.3f hist %.3f' % (d['ms_per_view'], k['blend_fwd']['avg_ms'], k['blend_bwd']['avg_ms'], k['preprocess_bwd']['avg_ms'], k['sort_onesweep_pass']['avg_ms'], k['sort_histogram']['avg_ms']))"
done
