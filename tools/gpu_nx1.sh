#!/bin/bash
set -u
N=${1:-2}; O=gpurun_out/${2:-nx1}; shift; shift
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-mlp --no-cpu-baseline "$@" > $O/run.json 2> $O/run.err
python - $O/run.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d value %.1fM e2e %.1fM ms/step %.3f" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
