#!/bin/bash
# The commands behind profiles/r03_exp_*.jsonl / r03_trace_*.txt (second half of round 2).  Each block was one gpurun call on one B200;
# output names as committed.  usage: bash scripts/r03_experiments.sh <block>   (blocks: tail_forms ablation lanes ahead alias in_step
# small_batch pool slab_queue nap trace)
S="python scripts/exp_env_sweep.py"
case "$1" in
tail_forms)  PASSES=4 $S cfg2,cfg5,mid,cfg3s "" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail_forms_4pass.jsonl ;;
ablation)    $S cfg3s "" "WANT_RESID=0" "WANT_Q=0" "WANT_Q=0 WANT_RESID=0" "VQB_RESID_REPLICAS=8" "VQB_RESID_REPLICAS=1" "VQB_TAIL_FORM=2 WANT_RESID=0" > gpurun_out/r03_exp_tail3_ablation.jsonl ;;
lanes)       $S cfg2,cfg5 "" "VQB_TAIL_LPF=8" "VQB_TAIL_LPF=2" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail3_lanes_per_frame.jsonl
             $S mid,cfg3s "" "VQB_TAIL_LPF=4" "VQB_TAIL_FORM=300" "VQB_TAIL_FORM=2" >> gpurun_out/r03_exp_tail3_lanes_per_frame.jsonl ;;
ahead)       PASSES=6 $S cfg3s,cfg2,cfg5,mid "VQB_TAIL_AHEAD=0" "VQB_TAIL_AHEAD=1" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail3_ahead_6pass.jsonl ;;
alias)       python scripts/exp_alias.py > gpurun_out/r03_exp_tail_address_alias.jsonl ;;
in_step)     export VQB_EXPERIMENTS=1
             for pass in 1 2; do for e in "VQB_TAIL_AHEAD=1" "VQB_TAIL_AHEAD=0" "VQB_TAIL_FORM=2"; do
               env $e python bench.py --no-e2e --no-cpu --no-train --no-sampler --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('%-20s' % '$e', 'search %.3f tail %.3f step %.3f' % (s['search'], s['tail'], d['ms_per_step']))"
             done; done > gpurun_out/r03_exp_tail_in_step_cfg3.txt ;;
small_batch) PASSES=3 $S cfg1,cfg5,cfg2 "" "VQB_TAIL_FORM=30" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail_small_batch.jsonl ;;
pool)        $S cfg2,cfg5,mid,cfg3s "" "VQB_TC_EVSM=-2" "" "VQB_TC_EVSM=-2" > gpurun_out/r03_exp_overflow_pool.jsonl ;;
slab_queue)  PASSES=3 $S cfg2,cfg5,cfg1 "" "VQB_TC_EPI=1" "PREC=tf32" "PREC=tf32 VQB_TC_EPI=1" > gpurun_out/r03_exp_slab_queue_epilogue.jsonl ;;
nap)         PASSES=3 $S cfg2,cfg5,mid,cfg3s "" "VQB_TC_SLEEP=40" "VQB_TC_SLEEP=100" "VQB_TC_SLEEP=200" "VQB_TC_SLEEP=400" "VQB_TC_SLEEP=100 PREC=tf32" "PREC=tf32" > gpurun_out/r03_exp_helper_nap.jsonl ;;
trace)       python scripts/build_trace.py   # (here, before the gpurun call: the debug library travels with the snapshot)
             python scripts/trace_tc.py cfg2 gpurun_out/r03_trace_cfg2.json > gpurun_out/r03_trace_cfg2_final.txt
             python scripts/trace_tc.py cfg5 gpurun_out/r03_trace_cfg5.json > gpurun_out/r03_trace_cfg5_final.txt
             VQB_EXPERIMENTS=1 VQB_TC_EPI=1 python scripts/trace_tc.py cfg2 gpurun_out/trace_cfg2_sq.json > gpurun_out/r03_trace_cfg2_slab_queue_epilogue.txt ;;
*) echo "usage: $0 tail_forms|ablation|lanes|ahead|alias|in_step|small_batch|pool|slab_queue|nap|trace" ;;
esac
