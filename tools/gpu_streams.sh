#!/bin/bash
set -u
O=gpurun_out/${1:-streams}
mkdir -p $O
for s in 2 4 6 8; do
timeout 300 python bench.py --steps 20 --warmup 5 --streams $s --no-cpu-baseline --no-mlp > $O/s$s.json 2> $O/s$s.err
python - $O/s$s.json $s <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("streams",sys.argv[2],"ms/view %.4f value %.1fM e2e %.1fM"%(d["ms_per_view"],d["value"]/1e6,d["e2e"]["value"]/1e6))
PY
done
