#!/bin/bash
# binning / forward stages bit-exact + kernel table
set -u
O=gpurun_out/${1:-q4}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_stages or edge_cases or golden_vectors or cpu_oracle or c4_6m or variant or wider_than" > $O/pytest.log 2>&1; tail -3 $O/pytest.log; grep -n "^E " $O/pytest.log | head -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mlp > $O/c2.json 2> $O/c2.err
python - $O <<'PY'
import json,sys
O=sys.argv[1]
try:
    d=json.loads(open("%s/c2.json"%O).read().strip().splitlines()[-1])
    print("c2 ms/view %.4f value %.1fM e2e %.1fM" % (d["ms_per_view"], d["value"]/1e6, d["e2e"]["value"]/1e6))
    for k,v in d["kernels"].items(): print("   %-28s x%-3d avg %.5f ms" % (k, v["launches"], v["avg_ms"]))
except Exception as e:
    print("ERR",e, open("%s/c2.err"%O).read()[-1500:])
PY
