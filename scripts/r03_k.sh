#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r03_k_tests.log 2>&1; tail -3 gpurun_out/r03_k_tests.log
timeout 300 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_s4.json > gpurun_out/r03_trace_cfg2_s4.txt 2>&1; tail -16 gpurun_out/r03_trace_cfg2_s4.txt | head -14 | cut -c330-620
