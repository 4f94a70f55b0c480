#!/bin/bash
# ncu launch list (per-launch device time, cold-cache and serialised: compare SHARES) of training steps run eagerly
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 3 --no-graph --frames 4 --no-cpu-baseline --no-gpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/r02_launches_step.csv; tail -2 gpurun_out/r02_ncu_plain.log | cut -c1-200
