#!/bin/bash
# N-GPU bench of the "ours" arm with and without the overlapped all-reduce: tools/gpu_nx.sh <N> <outdir>
set -u
N=${1:-2}
O=gpurun_out/${2:-nx}
mkdir -p $O
run() { # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-mlp --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err
  python - $O/$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%d value %.1fM e2e %.1fM ms/step %.3f" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]))
except Exception as e:
    print(sys.argv[2], "ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-mlp --no-cpu-baseline > $O/n1.json 2> $O/n1.err; python -c "
import json;d=json.loads(open('$O/n1.json').read().strip().splitlines()[-1]);print('n1 value %.1fM e2e %.1fM ms/step %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']))"
run overlap1 --overlap-allreduce 1
run overlap0 --overlap-allreduce 0
