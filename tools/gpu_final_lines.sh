#!/bin/bash
set -u
O=gpurun_out/final; mkdir -p $O
timeout 300 python bench.py --batched-backward 0 --batched-forward 0 --no-cpu-baseline --no-mlp > $O/bench_ours_c2_per_view.json 2> /dev/null
timeout 300 python bench.py --views 1 --streams 1 --no-cpu-baseline --no-mlp > $O/bench_ours_c2_one_view.json 2> /dev/null
timeout 300 python bench.py --config C3 --no-cpu-baseline --no-mlp > $O/bench_ours_c3.json 2> /dev/null
timeout 300 python bench.py --config C4 --no-cpu-baseline --no-mlp > $O/bench_ours_c4.json 2> /dev/null
timeout 300 python bench.py --train --no-cpu-baseline --no-mlp > $O/bench_ours_c2_train.json 2> /dev/null
python - <<'PY'
import json
for f in ("bench_ours_c2_per_view","bench_ours_c2_one_view","bench_ours_c3","bench_ours_c4","bench_ours_c2_train"):
    try:
        d=json.loads(open("gpurun_out/final/%s.json"%f).read().strip().splitlines()[-1])
        print("%-28s ms/view %.4f value %.1fM e2e %.1fM rast_only %s train %s"%(f,d["ms_per_view"],d["value"]/1e6,d["e2e"]["value"]/1e6,(d.get("breakdown") or {}).get("rasterizer_only_ms_per_view"), (d.get("train_step") or {}).get("ms_per_step")))
    except Exception as e: print(f,"ERR",e)
PY
