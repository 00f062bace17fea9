#!/bin/bash
export PASSES=6
timeout 600 python scripts/exp_env_sweep.py cfg3s,cfg2,cfg5,mid "VQB_TAIL_AHEAD=0" "VQB_TAIL_AHEAD=1" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_tail3_ahead2.jsonl 2> gpurun_out/r03_exp_tail3_ahead2.err
python - <<'PY'
import json
for l in open("gpurun_out/r03_exp_tail3_ahead2.jsonl"):
    d = json.loads(l); print(d["case"], "%-40s" % d["env"], "tail %.4f search %.4f step %.4f" % (d["tail"], d["search"], d["step_ms"]))
PY
