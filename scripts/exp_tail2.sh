#!/bin/bash
# stand-alone tail experiments: TMA-fed vs register-staged
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
for t in 1 0; do
  VQB_TAIL_TMA=$t timeout 120 python bench.py --steps 5 --no-e2e --no-cpu > gpurun_out/s3_b3_tma$t.log 2>&1
  VQB_TAIL_TMA=$t timeout 120 python bench.py --steps 5 --no-e2e --no-cpu --workload cfg2 > gpurun_out/s3_b2_tma$t.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tail -c 2 --csv --log-file gpurun_out/s3_ncu_tail.csv python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/s3_ncu_tail.log 2>&1
