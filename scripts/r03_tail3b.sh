#!/bin/bash
# what bounds tail3_kernel at D = 256: residual atomics on / off, quantized on / off; ncu --set full of one launch
timeout 300 python scripts/exp_env_sweep.py cfg3s "" "WANT_RESID=0" "WANT_Q=0" "WANT_Q=0 WANT_RESID=0" "VQB_RESID_REPLICAS=8" "VQB_RESID_REPLICAS=1" "VQB_TAIL_FORM=2 WANT_RESID=0" > gpurun_out/r03_exp_tail3b.jsonl 2> gpurun_out/r03_exp_tail3b.err
cat gpurun_out/r03_exp_tail3b.jsonl; tail -3 gpurun_out/r03_exp_tail3b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail3 -s 1 -c 1 -f -o gpurun_out/r03_tail3_cfg3 python scripts/profile_fwd.py 256 256 16384 8192 2 > gpurun_out/r03_ncu_tail3.log 2>&1; tail -2 gpurun_out/r03_ncu_tail3.log
