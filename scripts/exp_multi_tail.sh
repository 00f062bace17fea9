#!/bin/bash
# VERDICT r01 Next #1c: why is tail_tma_kernel slower whenever WORLD_SIZE > 1?  (run with gpurun --gpus 2)
out=gpurun_out/r02_exp_multi_tail.jsonl
: > $out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29500"
run() { echo "## $*" >> gpurun_out/r02_exp_multi_tail.err; timeout 240 "$@" >> $out 2>> gpurun_out/r02_exp_multi_tail.err; }
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
run python scripts/exp_multi_tail.py --variant plain
run $T --nproc-per-node 1 scripts/exp_multi_tail.py --variant plain
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant plain
run $T --nproc-per-node 1 scripts/exp_multi_tail.py --variant pg_only
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant pg_only
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant comm_idle
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant inplace
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant outofplace
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant sidestream
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant gloo_pg
NCCL_P2P_DISABLE=1 run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant inplace
run $T --nproc-per-node 2 scripts/exp_multi_tail.py --variant inplace --no-resid
run python scripts/exp_multi_tail.py --variant plain --no-resid
cat $out
