#!/bin/bash
# depth sort v2: correctness (stand-alone + pipeline stages) then kernel tables
set -u
O=gpurun_out/${1:-ds}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "depth_order or forward_stages or edge_cases or golden_vectors or cpu_oracle" > $O/pytest.log 2>&1; tail -3 $O/pytest.log; grep -n "^E " $O/pytest.log | head -20
timeout 300 python bench.py --views 1 --streams 1 --steps 10 --warmup 3 --no-cpu-baseline --no-mlp > $O/v1.json 2> $O/v1.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mlp > $O/c2.json 2> $O/c2.err
python - $O <<'PY'
import json,sys
O=sys.argv[1]
for f in ("v1","c2"):
    try:
        d=json.loads(open("%s/%s.json"%(O,f)).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e, open("%s/%s.err"%(O,f)).read()[-1500:]); continue
    print(f, "ms/view %.4f value %.1fM e2e %.1fM rast_only %.4f" % (d["ms_per_view"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["breakdown"]["rasterizer_only_ms_per_view"]))
    if f=="v1":
        for k,v in d["kernels"].items(): print("   %-28s x%-3d avg %.5f ms %s" % (k, v["launches"], v["avg_ms"], v.get("frac_of_hbm_peak", v.get("frac_of_fp32_issue_peak",""))))
PY
