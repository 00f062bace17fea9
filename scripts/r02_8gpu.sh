#!/bin/bash
# round-2 builder-run 8-GPU lines: cfg3 (strong + weak), cfg4, cfg5, and the 2-rank parity tests on the same box
t=${1:-r02}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${t}_topo_8gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${t}_tests_multi.log 2>&1; tail -2 gpurun_out/${t}_tests_multi.log
timeout 400 $T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > gpurun_out/${t}_bench_cfg3_8gpu.json 2> gpurun_out/${t}_bench_cfg3_8gpu.err; tail -c 300 gpurun_out/${t}_bench_cfg3_8gpu.err
timeout 400 $T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --workload cfg4 > gpurun_out/${t}_bench_cfg4_8gpu.json 2> gpurun_out/${t}_bench_cfg4_8gpu.err; tail -c 300 gpurun_out/${t}_bench_cfg4_8gpu.err
timeout 300 $T --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --workload cfg5 > gpurun_out/${t}_bench_cfg5_8gpu.json 2> gpurun_out/${t}_bench_cfg5_8gpu.err; tail -c 300 gpurun_out/${t}_bench_cfg5_8gpu.err
timeout 200 python bench.py --no-e2e --no-cpu --no-train --steps 8 > gpurun_out/${t}_bench_cfg3_1of8.json 2>/dev/null
ls -la gpurun_out/${t}_*
