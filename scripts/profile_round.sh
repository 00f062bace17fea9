#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command, full captures of the dominant kernels at cfg 3 and cfg 2 (tag = $1)
t=${1:-r03}
set -x
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/${t}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${t}_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-train > gpurun_out/${t}_ncu_launch.log 2>&1
python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/${t}_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tail3 -s 1 -c 1 -f -o gpurun_out/${t}_tail3_cfg3 python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/${t}_ncu_tail.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_search -s 1 -c 1 -f -o gpurun_out/${t}_tc_search_cfg3 python scripts/profile_fwd.py 1024 256 16384 8192 2 > gpurun_out/${t}_ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_search -s 1 -c 1 -f -o gpurun_out/${t}_tc_search_cfg2 python scripts/profile_fwd.py 64 64 16384 1024 3 > gpurun_out/${t}_ncu_tc2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail3 -s 1 -c 1 -f -o gpurun_out/${t}_tail3_cfg2 python scripts/profile_fwd.py 64 64 16384 1024 3 > gpurun_out/${t}_ncu_tail2.log 2>&1
