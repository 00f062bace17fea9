#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adversarial or lanes_per_frame or variants" > gpurun_out/r03_r_tests.log 2>&1; tail -5 gpurun_out/r03_r_tests.log
timeout 600 python scripts/stress_shapes.py 300 > gpurun_out/r03_stress.log 2>&1; tail -3 gpurun_out/r03_stress.log
