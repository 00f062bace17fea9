#!/bin/bash
# last check of a round: GPU tests, the default bench line, smoke (tag = $1)
t=${1:-r03}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${t}_tests_gpu.log 2>&1; tail -2 gpurun_out/${t}_tests_gpu.log
timeout 400 python bench.py > gpurun_out/${t}_bench_cfg3.json 2> gpurun_out/${t}_bench_cfg3.err; tail -c 200 gpurun_out/${t}_bench_cfg3.err
for w in cfg2 cfg5 cfg1; do timeout 300 python bench.py --workload $w > gpurun_out/${t}_bench_$w.json 2>/dev/null; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
