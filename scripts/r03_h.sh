#!/bin/bash
# grouped epilogue for small codebooks: parity, A/B against the per-thread stacks, timeline
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03_h_tests.log 2>&1; tail -5 gpurun_out/r03_h_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,cfg1,mid "" "VQB_TC_EPI=0" "PREC=tf32" "PREC=tf32 VQB_TC_EPI=0" > gpurun_out/r03_exp_epi.jsonl 2> gpurun_out/r03_exp_epi.err
cut -c1-330 gpurun_out/r03_exp_epi.jsonl; tail -3 gpurun_out/r03_exp_epi.err
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_q.json > gpurun_out/r03_trace_cfg2_q.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_q.txt | head -14
