#!/bin/bash
# same A/B at the full BASELINE config-3 size (B = 1024, N = 2^24 per GPU), where SCALE_r01 showed the +2.4 ms
out=gpurun_out/r02_exp_multi_tail_full.jsonl
: > $out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29500"
run() { echo "## $*" >> gpurun_out/r02_exp_multi_tail_full.err; timeout 300 "$@" >> $out 2>> gpurun_out/r02_exp_multi_tail_full.err; }
run python scripts/exp_multi_tail.py --variant plain --B 1024 --steps 12
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant plain --B 1024 --steps 12
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant pg_only --B 1024 --steps 12
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant comm_idle --B 1024 --steps 12
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant inplace --B 1024 --steps 12
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant sidestream --B 1024 --steps 12
run python scripts/exp_multi_tail.py --variant plain --B 1024 --steps 12
cat $out
