#!/bin/bash
# round-2 check on a 2-GPU lease: GPU tests (incl. the 2-GPU ones), single-GPU bench, 2-GPU bench (strong + weak)
t=${1:-r02a}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${t}_tests_gpu.log 2>&1; tail -3 gpurun_out/${t}_tests_gpu.log
timeout 400 python bench.py > gpurun_out/${t}_bench_cfg3.json 2> gpurun_out/${t}_bench_cfg3.err; tail -c 600 gpurun_out/${t}_bench_cfg3.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 > gpurun_out/${t}_bench_cfg3_2gpu.json 2> gpurun_out/${t}_bench_cfg3_2gpu.err; tail -c 600 gpurun_out/${t}_bench_cfg3_2gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${t}_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
