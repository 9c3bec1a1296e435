#!/bin/bash
set -u
mkdir -p gpurun_out/n2
O=gpurun_out/n2
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/ours_n$N.json 2> $O/ours_n$N.err
tail -3 $O/ours_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/n2/ours_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1fM e2e %.1fM ms/step %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), d["e2e"])
PY
