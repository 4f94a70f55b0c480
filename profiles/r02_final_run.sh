#!/bin/bash
# Round-2 evidence run (one B200): GPU test suite, the four BASELINE configs through bench.py, the ncu launch list of
# a training step and `ncu --set full` captures of the block-4 convolution kernels and the fold GEMMs.
cd "$GRAFT_REPO_ROOT"
ST=./boosting-neural-video-representation-via-online-structural-reparameteration_b200/onr_selftest
what=${1:-all}
if [ "$what" = all ] || [ "$what" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_final_pytest.log 2>&1; tail -3 gpurun_out/r02_final_pytest.log
fi
if [ "$what" = all ] || [ "$what" = bench ]; then
  for c in c1 c2 c3 c4; do
    timeout 400 python bench.py --config $c > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "bench $c rc=$?"
  done
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c1_reference_arm.json 2>/dev/null
  timeout 300 python bench.py --impl torch-gpu --steps 10 --warmup 3 > gpurun_out/r02_bench_c1_torch_gpu.json 2>/dev/null
fi
if [ "$what" = all ] || [ "$what" = ncu ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-graph --frames 4 --no-cpu-baseline --no-gpu-baseline"
  $CMD > gpurun_out/r02_ncu_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
  for op in fprop dgrad wgrad; do
    timeout 60 $ST $op l4 3 nocheck > /dev/null 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -o gpurun_out/r02_ncu_full_${op}_block4 $ST $op l4 1 nocheck > /dev/null 2>&1
  done
  python profiles/fold_probe.py 112 2800 2 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:fold_tc_gemm -s 6 -c 6 -o gpurun_out/r02_ncu_full_fold_gemm_L720 python profiles/fold_probe.py 112 2800 2 > /dev/null 2>&1
fi
ls gpurun_out | grep r02_ | head -40
