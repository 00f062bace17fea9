#!/bin/bash
PASSES=3 timeout 600 python scripts/exp_env_sweep.py cfg2,cfg5,mid,cfg3s "" "VQB_TC_SLEEP=40" "VQB_TC_SLEEP=100" "VQB_TC_SLEEP=200" "VQB_TC_SLEEP=400" "VQB_TC_SLEEP=100 PREC=tf32" "PREC=tf32" > gpurun_out/r03_exp_sleep.jsonl 2> gpurun_out/r03_exp_sleep.err
python - <<'PY'
import json
for l in open("gpurun_out/r03_exp_sleep.jsonl"):
    d = json.loads(l); print(d["case"], "%-28s" % d["env"], "search %.4f tail %.4f step %.4f" % (d["search"], d["tail"], d["step_ms"]), d.get("check","")[:40])
PY
tail -2 gpurun_out/r03_exp_sleep.err
