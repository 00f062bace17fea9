#!/bin/bash
# per-thread stacks with own slots + warp overflow pool in shared memory
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03_j_tests.log 2>&1; tail -3 gpurun_out/r03_j_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,mid,cfg3s "" "VQB_TC_EVSM=-2" "" "VQB_TC_EVSM=-2" > gpurun_out/r03_exp_pool.jsonl 2> gpurun_out/r03_exp_pool.err
cut -c1-330 gpurun_out/r03_exp_pool.jsonl; tail -3 gpurun_out/r03_exp_pool.err
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_s3.json > gpurun_out/r03_trace_cfg2_s3.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_s3.txt | head -14 | cut -c1-420
