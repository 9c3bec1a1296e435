#!/bin/bash
set -u
mkdir -p gpurun_out/quick
O=gpurun_out/quick
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contract.py -m gpu -q -x -k "forward_stages or forward_backward_vs or golden_vectors or cpu_oracle or edge_cases or c4_6m or variant or accumulate or sync_free" > $O/pytest.log 2>&1; tail -3 $O/pytest.log; grep -n "^E " $O/pytest.log | head
for mk in 0 1; do
GSR_BLEND_MASKS=$mk timeout 300 python bench.py --views 1 --streams 1 --steps 20 --warmup 3 --no-cpu-baseline --no-mlp > $O/v1_mk$mk.json 2> $O/v1_mk$mk.err
GSR_BLEND_MASKS=$mk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mlp > $O/c2_mk$mk.json 2> $O/c2_mk$mk.err
done
python - <<'PY'
import json
for f in ("v1_mk0","v1_mk1","c2_mk0","c2_mk1"):
    try:
        d=json.loads(open("gpurun_out/quick/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e, open("gpurun_out/quick/%s.err"%f).read()[-800:]); continue
    k=d["kernels"]
    print(f, "ms/view %.4f value %.1fM e2e %.1fM | blend_fwd %.4f blend_bwd %.4f" % (d["ms_per_view"], d["value"]/1e6, d["e2e"]["value"]/1e6, k["blend_fwd"]["avg_ms"], k["blend_bwd"]["avg_ms"]))
PY
