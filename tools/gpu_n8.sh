#!/bin/bash
set -u
mkdir -p gpurun_out/n8
O=gpurun_out/n8
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/ours_n$N.json 2> $O/ours_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --train > $O/ours_train_n$N.json 2> $O/ours_train_n$N.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-mlp --no-cpu-baseline > $O/ours_n1.json 2> $O/ours_n1.err
python - <<PY
import json
for f in ("ours_n$N","ours_train_n$N","ours_n1"):
    try:
        d=json.loads(open("gpurun_out/n8/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.1fM e2e %.1fM ms/step %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), d.get("train_step",{}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/n8/%s.err"%f).read()[-500:])
PY
