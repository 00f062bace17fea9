#!/bin/bash
# tail experiments: bench with parts of the fused tail switched off (VQB_TAIL_DBG bits, see vqb_tc.cu)
for dbg in ${DBGS:-0 32}; do
  VQB_TAIL_DBG=$dbg timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train > gpurun_out/s2_dbg_$dbg.log 2>&1
done
if [ -n "$NCU" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tc_search -c 1 --csv --log-file gpurun_out/s2_ncu_dram.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/s2_ncu_dram.log 2>&1
fi
if [ -n "$TESTS" ]; then timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; tail -3 gpurun_out/s2_tests.log; fi
