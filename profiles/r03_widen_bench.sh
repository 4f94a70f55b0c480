#!/bin/bash
# Round-3 (second session of round 2) bench lines of the widened rows on one B200.
# Usage (on the GPU box, repo root): bash profiles/r03_widen_bench.sh
mkdir -p gpurun_out
run() { name=$1; shift; timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r03_bench_$name.json 2> gpurun_out/r03_bench_$name.err || echo "$name rc=$?"; }
run c1_dbb --branch_type DBB
run c1_ecb --branch_type ECB
run c1_gelu --act gelu
run c1_finetune --finetune_prune 0.4
for f in gpurun_out/r03_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    g=d.get('gpu_library_baseline') or {}
    print(sys.argv[1].split('/')[-1], 'value %.1f e2e %.1f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'torch-gpu', g.get('value'), 'psnr', d['last_step']['psnr'], 'ref_api', (d.get('e2e_reference_api') or {}).get('value'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
tail -n 5 gpurun_out/r03_bench_*.err
