#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_tests.log 2>&1; tail -3 gpurun_out/s5_tests.log
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['losses']['embedding'])
"; }
for r in 1 2 4 8; do VQB_RESID_REPLICAS=$r VQB_TAIL_VARIANT=1 timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train 2>/dev/null | show "v1 rep$r"; done
for v in 0 2; do VQB_TAIL_VARIANT=$v timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train 2>/dev/null | show "v$v rep4"; done
