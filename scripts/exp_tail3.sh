#!/bin/bash
VQB_TAIL_VARIANT=2 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests_v2.log 2>&1; tail -3 gpurun_out/s4_tests_v2.log
for v in 0 1 2; do VQB_TAIL_VARIANT=$v python scripts/bench_tail_only.py 8192 6 2>&1 | tail -1; done > gpurun_out/s4_var.log
