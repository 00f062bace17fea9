#!/bin/bash
# tail3_kernel: 4 / 8 compute warps x red.v4 / bulk reductions
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or variants or random_shapes" > gpurun_out/r03_tail3c_tests.log 2>&1; tail -3 gpurun_out/r03_tail3c_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg3s,mid,cfg2 "VQB_TAIL_FORM=300" "VQB_TAIL_FORM=301" "VQB_TAIL_FORM=302" "VQB_TAIL_FORM=303" "VQB_TAIL_FORM=303 VQB_RESID_REPLICAS=1" "VQB_TAIL_FORM=303 VQB_RESID_REPLICAS=4" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail3c.jsonl 2> gpurun_out/r03_exp_tail3c.err
cat gpurun_out/r03_exp_tail3c.jsonl; tail -3 gpurun_out/r03_exp_tail3c.err
