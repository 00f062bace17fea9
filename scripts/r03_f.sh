#!/bin/bash
# tail3 option A/B (list prefetch, |x|^2 folded into the pair scoring) and the accurate hand-off timeline of the search kernel
timeout 300 python scripts/exp_env_sweep.py mid,cfg3s,cfg2 "" "VQB_TAIL_OPTS=0" "VQB_TAIL_OPTS=1" "VQB_TAIL_OPTS=2" > gpurun_out/r03_exp_opts.jsonl 2> gpurun_out/r03_exp_opts.err
cut -c1-330 gpurun_out/r03_exp_opts.jsonl; tail -3 gpurun_out/r03_exp_opts.err
timeout 300 python scripts/trace_tc.py cfg2 > gpurun_out/r03_trace_cfg2.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2.txt
timeout 300 python scripts/trace_tc.py cfg5 > gpurun_out/r03_trace_cfg5.txt 2>&1; tail -3 gpurun_out/r03_trace_cfg5.txt
