#!/bin/bash
set -u
O=gpurun_out/${1:-bf}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched_forward or batched_backward or several_streams" > $O/pytest.log 2>&1; tail -3 $O/pytest.log; grep -n "^E " $O/pytest.log | head -20
run() { name=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mlp "$@" > $O/$name.json 2> $O/$name.err
  python - $O/$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d["kernels"]
    print(sys.argv[2],"ms/view %.4f value %.1fM e2e %.1fM"%(d["ms_per_view"],d["value"]/1e6,d["e2e"]["value"]/1e6), {n:k[n]["avg_ms"] for n in k if n.startswith("preprocess")})
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
run bf0 --batched-forward 0
run bf1 --batched-forward 1
run bf1_c3 --batched-forward 1 --config C3
