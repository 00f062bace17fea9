#!/bin/bash
# final single-GPU measurements of the round: tests, bench lines, reference arm, smoke, experiments, timelines, ncu evidence (tag = $1)
t=${1:-r03}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${t}_tests_gpu.log 2>&1; tail -2 gpurun_out/${t}_tests_gpu.log
timeout 400 python bench.py > gpurun_out/${t}_bench_cfg3.json 2> gpurun_out/${t}_bench_cfg3.err
for w in cfg2 cfg1 cfg5 cfg4; do timeout 300 python bench.py --workload $w > gpurun_out/${t}_bench_$w.json 2>/dev/null; done
timeout 300 python bench.py --workload cfg2 --precision tf32 > gpurun_out/${t}_bench_cfg2_tf32.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${t}_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
timeout 300 python scripts/exp_precisions.py --with-fp32 > gpurun_out/${t}_exp_precisions.jsonl 2>/dev/null
timeout 300 python scripts/exp_env_sweep.py cfg3s,mid,cfg2,cfg5 "" "VQB_TAIL_FORM=2" "VQB_TAIL_FORM=0" "" "VQB_TAIL_FORM=2" "VQB_TAIL_FORM=0" > gpurun_out/${t}_exp_tail_forms.jsonl 2>/dev/null
timeout 200 python scripts/trace_tc.py cfg2 gpurun_out/${t}_trace_cfg2.json > gpurun_out/${t}_trace_cfg2.txt 2>&1
timeout 200 python scripts/trace_tc.py cfg5 gpurun_out/${t}_trace_cfg5.json > gpurun_out/${t}_trace_cfg5.txt 2>&1
bash scripts/profile_round.sh $t > gpurun_out/${t}_profile.log 2>&1
ls -la gpurun_out/${t}_* | head -50
