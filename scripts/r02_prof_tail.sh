#!/bin/bash
# ncu --set full of the tail kernels (old form and tail2) at a 2^22-frame cfg-3 slice
set -x
export VQB_EXPERIMENTS=1
VQB_TAIL_FORM=2 python scripts/profile_fwd.py 256 256 16384 8192 2 > gpurun_out/r02_prof_plain.log 2>&1 || exit 1
VQB_TAIL_FORM=2 ncu --set full --clock-control none --import-source on -k regex:tail2 -s 1 -c 1 -f -o gpurun_out/r02_tail2_tf16 python scripts/profile_fwd.py 256 256 16384 8192 2 > gpurun_out/r02_ncu_tail2.log 2>&1
VQB_TAIL_FORM=232 ncu --set full --clock-control none --import-source on -k regex:tail2 -s 1 -c 1 -f -o gpurun_out/r02_tail2_tf32 python scripts/profile_fwd.py 256 256 16384 8192 2 > gpurun_out/r02_ncu_tail2b.log 2>&1
