#!/bin/bash
# slab-queue epilogue: the whole parity file with it forced on, A/B against the per-thread stacks, timeline
VQB_EXPERIMENTS=1 VQB_TC_EPI=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03_y_tests.log 2>&1; tail -3 gpurun_out/r03_y_tests.log
PASSES=3 timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,cfg1 "" "VQB_TC_EPI=1" "PREC=tf32" "PREC=tf32 VQB_TC_EPI=1" > gpurun_out/r03_exp_sq.jsonl 2> gpurun_out/r03_exp_sq.err
cut -c1-215 gpurun_out/r03_exp_sq.jsonl; tail -3 gpurun_out/r03_exp_sq.err
VQB_EXPERIMENTS=1 VQB_TC_EPI=1 timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_sq.json > gpurun_out/r03_trace_cfg2_sq.txt 2>&1; grep "tile  44\|tile  45 \|tile  46 \|tile  47 \|mean" gpurun_out/r03_trace_cfg2_sq.txt | head -5 | sed 's/ slab0 in/\n   slab0 in/; s/ resolve/\n   resolve/'
