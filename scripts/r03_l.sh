#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03_l_tests.log 2>&1; tail -3 gpurun_out/r03_l_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,cfg1,mid "" "" "PREC=tf32" > gpurun_out/r03_exp_own.jsonl 2> gpurun_out/r03_exp_own.err
cut -c1-200 gpurun_out/r03_exp_own.jsonl; tail -3 gpurun_out/r03_exp_own.err
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_s5.json > gpurun_out/r03_trace_cfg2_s5.txt 2>&1; grep "tile  44\|tile  45 \|mean" gpurun_out/r03_trace_cfg2_s5.txt | head -3 | sed 's/ slab0 in/\n   slab0 in/; s/ resolve/\n   resolve/'
