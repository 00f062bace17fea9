#!/bin/bash
# final single-GPU measurements of the round: tests, bench lines, reference arm, smoke, ncu evidence (tag = $1, default r01)
t=${1:-r01}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${t}_tests_gpu.log 2>&1; tail -2 gpurun_out/${t}_tests_gpu.log
timeout 400 python bench.py > gpurun_out/${t}_bench_cfg3.json 2> gpurun_out/${t}_bench_cfg3.err
for w in cfg2 cfg1 cfg5 cfg4; do timeout 300 python bench.py --workload $w > gpurun_out/${t}_bench_$w.json 2>/dev/null; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${t}_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
bash scripts/profile_round.sh $t > gpurun_out/${t}_profile.log 2>&1
ls -la gpurun_out/${t}_*
