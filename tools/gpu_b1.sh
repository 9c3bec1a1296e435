#!/bin/bash
# one default bench line + kernel table: tools/gpu_b1.sh <outdir> [bench args]
set -u
O=gpurun_out/${1:-b1}; shift
mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mlp "$@" > $O/c2.json 2> $O/c2.err
python - $O <<'PY'
import json,sys
O=sys.argv[1]
try:
    d=json.loads(open("%s/c2.json"%O).read().strip().splitlines()[-1])
    print("ms/view %.4f value %.1fM e2e %.1fM" % (d["ms_per_view"], d["value"]/1e6, d["e2e"]["value"]/1e6))
    for k,v in d["kernels"].items(): print("   %-28s x%-3d avg %.5f ms" % (k, v["launches"], v["avg_ms"]))
except Exception as e:
    print("ERR",e, open("%s/c2.err"%O).read()[-1500:])
PY
