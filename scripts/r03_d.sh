#!/bin/bash
# tail3 defaults in the full-size step, its ncu capture (current build), and the hand-off timeline of the search kernel at small D
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or variants or golden" > gpurun_out/r03_d_tests.log 2>&1; tail -3 gpurun_out/r03_d_tests.log
timeout 400 python bench.py --no-cpu --no-e2e --no-train > gpurun_out/r03_d_bench_cfg3.json 2> gpurun_out/r03_d_bench_cfg3.err; tail -c 300 gpurun_out/r03_d_bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r03_d_bench_cfg3.json')); print(d['ms_per_step'], d['stage_ms_per_step'], d.get('roofline_tail'))"
timeout 300 python scripts/trace_tc.py cfg2 > gpurun_out/r03_trace_cfg2.txt 2>&1; tail -32 gpurun_out/r03_trace_cfg2.txt
timeout 300 python scripts/trace_tc.py cfg5 > gpurun_out/r03_trace_cfg5.txt 2>&1; tail -3 gpurun_out/r03_trace_cfg5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail3 -s 1 -c 1 -f -o gpurun_out/r03_tail3_cfg3 python scripts/profile_fwd.py 256 256 16384 8192 2 > gpurun_out/r03_ncu_tail3.log 2>&1; tail -1 gpurun_out/r03_ncu_tail3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail2 -s 1 -c 1 -f -o gpurun_out/r03_tail2_cfg2 python scripts/profile_fwd.py 64 64 16384 1024 2 > gpurun_out/r03_ncu_tail2.log 2>&1; tail -1 gpurun_out/r03_ncu_tail2.log
