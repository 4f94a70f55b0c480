#!/bin/bash
# Launch list of one bench step (our kernels only), per B200_PROFILING.md: run the plain command first, then the same
# command under ncu.  Usage (on the GPU box): bash profiles/launch_list.sh <tag>   -> gpurun_out/launches_<tag>.csv
set -u
tag=${1:-r01}
cmd="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --frames 8"
mkdir -p gpurun_out
timeout 120 $cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/plain_$tag.log; exit 1; }
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --kernel-name-base demangled -k regex:onr:: -c 700 --csv --log-file gpurun_out/launches_$tag.csv $cmd \
    > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
