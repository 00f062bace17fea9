#!/bin/bash
# tail3_kernel: look-ahead option x lanes per frame per shape, three round-robin passes, smallest step time per environment
export PASSES=3
timeout 600 python scripts/exp_env_sweep.py cfg2,cfg5 "VQB_TAIL_AHEAD=0" "VQB_TAIL_AHEAD=1" "VQB_TAIL_AHEAD=0 VQB_TAIL_LPF=8" "VQB_TAIL_AHEAD=1 VQB_TAIL_LPF=8" "VQB_TAIL_AHEAD=0 VQB_TAIL_LPF=2" "VQB_TAIL_AHEAD=1 VQB_TAIL_LPF=2" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail3_ahead.jsonl 2> gpurun_out/r03_exp_tail3_ahead.err
timeout 600 python scripts/exp_env_sweep.py mid,cfg3s "VQB_TAIL_AHEAD=0" "VQB_TAIL_AHEAD=1" "VQB_TAIL_AHEAD=0 VQB_TAIL_LPF=4" "VQB_TAIL_AHEAD=1 VQB_TAIL_LPF=4" "VQB_TAIL_FORM=2" >> gpurun_out/r03_exp_tail3_ahead.jsonl 2>> gpurun_out/r03_exp_tail3_ahead.err
python - <<'PY'
import json
for l in open("gpurun_out/r03_exp_tail3_ahead.jsonl"):
    d = json.loads(l); print(d["case"], "%-40s" % d["env"], "tail %.4f step %.4f" % (d["tail"], d["step_ms"]))
PY
tail -3 gpurun_out/r03_exp_tail3_ahead.err
