#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "oracle_parity or golden or lanes_per_frame or random_shapes or variants" > gpurun_out/r03_ad_tests.log 2>&1; tail -3 gpurun_out/r03_ad_tests.log
PASSES=4 timeout 300 python scripts/exp_env_sweep.py cfg2,cfg5,mid,cfg1 "" "VQB_TAIL_FORM=2" > gpurun_out/r03_exp_unpipe.jsonl 2> gpurun_out/r03_exp_unpipe.err
python - <<'PY'
import json
for l in open("gpurun_out/r03_exp_unpipe.jsonl"):
    d = json.loads(l); print(d["case"], "%-20s" % d["env"], "tail %.4f search %.4f step %.4f" % (d["tail"], d["search"], d["step_ms"]))
PY
