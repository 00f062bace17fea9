#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command, full captures of the two dominant kernels
set -x
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/r01b_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/r01b_ncu_launch.log 2>&1
python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/r01b_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tail_tma -s 1 -c 1 -f -o gpurun_out/r01b_tail_cfg3 python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/r01b_ncu_tail.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_search -s 1 -c 1 -f -o gpurun_out/r01b_tc_search_cfg3 python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/r01b_ncu_tc.log 2>&1
