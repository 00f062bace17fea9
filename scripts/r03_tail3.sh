#!/bin/bash
# tail3_kernel: parity tests that exercise the tail, then A/B against tail2 / the round-1 tail at the BASELINE shapes
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or golden or variants or random_shapes or nan_and or duplicate or host_buffer or graph" > gpurun_out/r03_tail3_tests.log 2>&1; tail -5 gpurun_out/r03_tail3_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg3s,cfg2,cfg5,mid "" "VQB_TAIL_FORM=2" "VQB_TAIL_FORM=0" > gpurun_out/r03_exp_tail3.jsonl 2> gpurun_out/r03_exp_tail3.err
cat gpurun_out/r03_exp_tail3.jsonl; tail -3 gpurun_out/r03_exp_tail3.err
