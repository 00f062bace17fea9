#!/bin/bash
export VQB_EXPERIMENTS=1 VQB_TC_EPI=1
python scripts/profile_fwd.py 64 64 16384 1024 3 > gpurun_out/r03_z_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_search -s 1 -c 1 -f -o gpurun_out/r03_tc_search_cfg2_slabqueue python scripts/profile_fwd.py 64 64 16384 1024 3 > gpurun_out/r03_z_ncu.log 2>&1; tail -1 gpurun_out/r03_z_ncu.log
