#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03_ae_tests.log 2>&1; tail -3 gpurun_out/r03_ae_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03_ae_smoke.log 2>&1; tail -1 gpurun_out/r03_ae_smoke.log
PASSES=3 timeout 300 python scripts/exp_env_sweep.py cfg1,cfg5,cfg2 "" "VQB_TAIL_FORM=30" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_small_batch.jsonl 2> gpurun_out/r03_exp_small_batch.err
python - <<'PY'
import json
for l in open("gpurun_out/r03_exp_small_batch.jsonl"):
    d = json.loads(l); print(d["case"], "%-20s" % d["env"], "tail %.4f search %.4f step %.4f" % (d["tail"], d["search"], d["step_ms"]))
PY
timeout 200 python bench.py --workload cfg1 --no-e2e --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg1 bench ms/step %.4f' % d['ms_per_step'], d['stage_ms_per_step'])"
