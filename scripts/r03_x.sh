#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03_x_tests.log 2>&1; tail -3 gpurun_out/r03_x_tests.log
PASSES=3 timeout 300 python scripts/exp_env_sweep.py cfg1,cfg2,cfg5 "" > gpurun_out/r03_exp_memset.jsonl 2> gpurun_out/r03_exp_memset.err; cut -c1-200 gpurun_out/r03_exp_memset.jsonl
for w in cfg1 cfg2 cfg5; do timeout 200 python bench.py --workload $w --no-e2e --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', 'ms/step %.4f' % d['ms_per_step'], 'launches', d.get('gpu_launches'))"; done
