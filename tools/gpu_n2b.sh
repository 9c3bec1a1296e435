#!/bin/bash
set -u
mkdir -p gpurun_out/n2
O=gpurun_out/n2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/gpu_adam_pipeline_check.py > $O/adam_pipeline.log 2>&1; echo "rc=$?" >> $O/adam_pipeline.log
grep -E "PIPELINE|rc=|Error" $O/adam_pipeline.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --train --no-mlp > $O/ours_train_n2.json 2> $O/ours_train_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/n2/ours_train_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1fM e2e %.1fM ms/step %.3f train %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["train_step"]["ms_per_step"]))
PY
