#!/bin/bash
# per-thread stacks with one atomic per event in the resolution; grouped epilogue as the opt-in
timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,mid,cfg3s "" "VQB_TC_EPI=1" "" "VQB_TC_EPI=1" > gpurun_out/r03_exp_epi2.jsonl 2> gpurun_out/r03_exp_epi2.err
cut -c1-330 gpurun_out/r03_exp_epi2.jsonl; tail -3 gpurun_out/r03_exp_epi2.err
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_s2.json > gpurun_out/r03_trace_cfg2_s2.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_s2.txt | head -14 | cut -c1-420
