#!/bin/bash
# correctness of the convolution kernels on every self-test shape (tcgen05 vs SIMT), then block-4 timing
cd "$GRAFT_REPO_ROOT"
ST=./boosting-neural-video-representation-via-online-structural-reparameteration_b200/onr_selftest
fail=0
for op in fprop infer dgrad wgrad; do for sh in tiny l0 l1 l2s b2 u3 wide l3; do
  timeout 60 $ST $op $sh 0 > gpurun_out/conv_check_tmp.log 2>&1 || { echo "FAIL $op $sh rc=$?"; tail -3 gpurun_out/conv_check_tmp.log; fail=1; }
done; done
echo "conv correctness fail=$fail"
for op in fprop infer dgrad wgrad; do timeout 60 $ST $op l4 20 nocheck 2>&1 | grep -E "time"; done
