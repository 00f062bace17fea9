#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or golden or variants or random_shapes or fused_tail or million" > gpurun_out/r03_q_tests.log 2>&1; tail -3 gpurun_out/r03_q_tests.log
PASSES=3 timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,mid,cfg3s "" "PREC=tf32" > gpurun_out/r03_exp_pretest.jsonl 2> gpurun_out/r03_exp_pretest.err
cut -c1-200 gpurun_out/r03_exp_pretest.jsonl; tail -3 gpurun_out/r03_exp_pretest.err
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_s6.json > gpurun_out/r03_trace_cfg2_s6.txt 2>&1; grep "tile  44\|tile  45 \|mean" gpurun_out/r03_trace_cfg2_s6.txt | head -3 | sed 's/ slab0 in/\n   slab0 in/; s/ resolve/\n   resolve/'
