#!/bin/bash
# round-3 session start: GPU tests of the restored tree + small-shape experiments (fused tail, 1-CTA mode, ring epilogue)
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r03_tests_gpu.log 2>&1; tail -3 gpurun_out/r03_tests_gpu.log
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,cfg1 "" "VQB_TC_TAIL=1" "VQB_TC_TAIL=1 VQB_TC_MODE=1" "VQB_TC_MODE=1" "VQB_TC_EPI=1" "PREC=tf32" > gpurun_out/r03_exp_small.jsonl 2> gpurun_out/r03_exp_small.err
cat gpurun_out/r03_exp_small.jsonl; tail -3 gpurun_out/r03_exp_small.err
