#!/bin/bash
# final single-GPU measurements of the round: tests, bench lines, reference arm, ncu evidence
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r01_tests_gpu.log 2>&1; tail -2 gpurun_out/r01_tests_gpu.log
timeout 400 python bench.py > gpurun_out/r01_bench_cfg3.json 2> gpurun_out/r01_bench_cfg3.err
timeout 300 python bench.py --workload cfg2 > gpurun_out/r01_bench_cfg2.json 2>/dev/null
timeout 300 python bench.py --workload cfg1 > gpurun_out/r01_bench_cfg1.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.log 2>&1; tail -1 gpurun_out/r01_smoke.log
bash scripts/profile_round.sh > gpurun_out/r01b_profile.log 2>&1
