#!/bin/bash
set -u
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
rm -f gpurun_out/grad_report.jsonl gpurun_out/mlp_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/grad_report.jsonl gpurun_out/mlp_report.jsonl $O/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench_ours_c2.json 2> $O/bench_ours_c2.err
timeout 500 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref_c2.json 2> $O/bench_ref_c2.err
tail -6 $O/pytest.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench_ours_c2.json; tail -3 $O/bench_ours_c2.err; python - <<'PY'
import json
for f in ("bench_ours_c2","bench_ref_c2"):
    try:
        d=json.loads(open("gpurun_out/r2k/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"]["value"], d.get("deform_network"))
    except Exception as e: print(f, "ERR", e)
PY
