#!/bin/bash
# Round-end verification on the GPU box: GPU tests, smoke, default bench (with CPU baseline), reference arm,
# launch list of one step, ncu --set full of the streaming head kernels.  Usage: bash profiles/final_run.sh <tag>
tag=${1:-r01final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
ONR_BENCH_WATCHDOG_S=300 timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
bash profiles/launch_list.sh $tag
cmd="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --frames 8"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'head_(bwd|fwd)_stream' --launch-skip 8 -c 2 -f -o gpurun_out/${tag}_head_stream $cmd > gpurun_out/ncu_head_$tag.log 2>&1
echo "ncu head rc=$?"
