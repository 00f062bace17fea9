#!/bin/bash
# tail3 with compile-time options; timelines of the ring epilogue and of the tf32 variant
timeout 300 python scripts/exp_env_sweep.py mid,cfg3s,cfg2,cfg5 "" "VQB_TAIL_LPF=4" "VQB_TAIL_LPF=8" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_t3final.jsonl 2> gpurun_out/r03_exp_t3final.err
cut -c1-330 gpurun_out/r03_exp_t3final.jsonl; tail -3 gpurun_out/r03_exp_t3final.err
VQB_EXPERIMENTS=1 VQB_TC_EPI=1 timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_ring.json > gpurun_out/r03_trace_cfg2_ring.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_ring.txt
PREC=tf32 timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_tf32.json > gpurun_out/r03_trace_cfg2_tf32.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_tf32.txt | head -13
