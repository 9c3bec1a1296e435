#!/bin/bash
set -u
mkdir -p gpurun_out/quick
O=gpurun_out/quick
timeout 600 python -m pytest tests/test_gpu_contract.py -m gpu -q -x -k "sync_free or prefiltered or debug_mode or interleaved" > $O/pytest.log 2>&1; tail -3 $O/pytest.log; grep -n "^E " $O/pytest.log | head
for sf in 0 1; do
timeout 300 python bench.py --views 1 --streams 1 --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --sync-free $sf > $O/v1_sf$sf.json 2> $O/v1_sf$sf.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mlp --sync-free $sf > $O/c2_sf$sf.json 2> $O/c2_sf$sf.err
timeout 300 python bench.py --steps 10 --warmup 3 --streams 1 --no-cpu-baseline --no-mlp --sync-free $sf > $O/c2s1_sf$sf.json 2> $O/c2s1_sf$sf.err
done
python - <<'PY'
import json
for f in ("v1_sf0","v1_sf1","c2_sf0","c2_sf1","c2s1_sf0","c2s1_sf1"):
    try:
        d=json.loads(open("gpurun_out/quick/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e, open("gpurun_out/quick/%s.err"%f).read()[-800:]); continue
    print(f, "ms/view %.4f value %.1fM e2e %.1fM rast_only %.4f" % (d["ms_per_view"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["breakdown"]["rasterizer_only_ms_per_view"]))
PY
