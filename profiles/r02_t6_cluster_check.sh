#!/bin/bash
# correctness of the clustered (multicast) kernels on small / ragged shapes, then timing on the block-4 shape
cd "$GRAFT_REPO_ROOT"
ST=./boosting-neural-video-representation-via-online-structural-reparameteration_b200/onr_selftest
fail=0
for op in fprop infer dgrad; do for sh in tiny l0 l1 l2s b2 u3 wide l3; do
  ONR_CONV_CLUSTER=2 timeout 60 $ST $op $sh 0 > gpurun_out/t6_tmp.log 2>&1 || { echo "FAIL cluster2 $op $sh rc=$?"; tail -3 gpurun_out/t6_tmp.log; fail=1; }
done; done
for sh in tiny l0 l1 l2s b2 u3 wide l3; do
  ONR_WGRAD_CLUSTER=3 timeout 60 $ST wgrad $sh 0 > gpurun_out/t6_tmp.log 2>&1 || { echo "FAIL cluster3 wgrad $sh rc=$?"; tail -3 gpurun_out/t6_tmp.log; fail=1; }
done
echo "cluster correctness fail=$fail"
for op in fprop dgrad; do for c in 1 2; do echo "== $op l4 cluster=$c"; ONR_CONV_CLUSTER=$c timeout 60 $ST $op l4 20 nocheck 2>&1 | tail -2; done; done
for c in 1 3; do echo "== wgrad l4 cluster=$c"; ONR_WGRAD_CLUSTER=$c timeout 60 $ST wgrad l4 20 nocheck 2>&1 | tail -2; done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "erb_fold" --timeout 120 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_t6_c1.json 2> gpurun_out/r2_t6_c1.err; tail -c 200 gpurun_out/r2_t6_c1.json
ONR_CONV_CLUSTER=1 ONR_WGRAD_CLUSTER=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_t6_c1_nocluster.json 2> /dev/null
