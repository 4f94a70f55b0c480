#!/bin/bash
# Runs every tcgen05-vs-SIMT cross-check case in its own process under a timeout.
# Usage: ./run_selftest.sh [out_log]
PKG="boosting-neural-video-representation-via-online-structural-reparameteration_b200"
LOG=${1:-gpurun_out/selftest.log}
mkdir -p "$(dirname "$LOG")"
: > "$LOG"
fail=0
run() {
  echo "### $*" >> "$LOG"
  timeout 60 "./$PKG/onr_selftest" "$@" >> "$LOG" 2>&1
  rc=$?
  echo "### rc=$rc" >> "$LOG"
  if [ $rc -ne 0 ]; then fail=1; fi
}
for op in fprop dgrad wgrad; do
  for shape in tiny l0 l1 l2s b2 u3 wide xl; do
    run $op $shape 0
  done
done
run infer l1 0
run fprop l3 5
run dgrad l3 5
run wgrad l3 5
run fprop l4 10 nocheck
run infer l4 10 nocheck
run dgrad l4 10 nocheck
run wgrad l4 10 nocheck
grep -E "PASS|FAIL|rc=|time|error" "$LOG"
exit $fail
