#!/bin/bash
# tail3 with 8 / 4 / 2 lanes per frame + shortlist prefetch; hand-off timeline of the search kernel (non-atomic trace)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or variants or golden or random_shapes" > gpurun_out/r03_e_tests.log 2>&1; tail -3 gpurun_out/r03_e_tests.log
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5 "" "VQB_TAIL_LPF=8" "VQB_TAIL_LPF=2" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_lpf.jsonl 2> gpurun_out/r03_exp_lpf.err
timeout 300 python scripts/exp_env_sweep.py mid,cfg3s "" "VQB_TAIL_LPF=4" "VQB_TAIL_FORM=300" "VQB_TAIL_FORM=2" >> gpurun_out/r03_exp_lpf.jsonl 2>> gpurun_out/r03_exp_lpf.err
cut -c1-330 gpurun_out/r03_exp_lpf.jsonl; tail -3 gpurun_out/r03_exp_lpf.err
timeout 300 python scripts/trace_tc.py cfg2 > gpurun_out/r03_trace_cfg2.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2.txt
timeout 300 python scripts/trace_tc.py cfg5 > gpurun_out/r03_trace_cfg5.txt 2>&1; tail -3 gpurun_out/r03_trace_cfg5.txt
