#!/bin/bash
# usage: bash profiles/r02_scale.sh N "c1 c2 ..." [steps]   -> gpurun_out/r02_scale_<cfg>_n<N>.json  (one JSON line each)
N=$1; CFGS=$2; STEPS=${3:-40}
cd "$GRAFT_REPO_ROOT"
for c in $CFGS; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --config $c --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline --no-gpu-baseline \
    > gpurun_out/r02_scale_${c}_n${N}.json 2> gpurun_out/r02_scale_${c}_n${N}.err
  echo "rc=$? $c N=$N"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_${c}_n${N}.json").read().strip().splitlines()[-1])
    print("   ", round(d["value"],1), "fps", round(d["ms_per_step"],4), "ms/step  e2e", round(d["e2e"]["value"],1), d["config"].get("gradient_exchange"))
except Exception as e:
    print("    no json:", e)
PY
done
