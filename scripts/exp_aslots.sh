#!/bin/bash
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['losses']['embedding'])
"; }
for a in 8 6 8 6; do VQB_TC_ASLOTS=$a timeout 120 python bench.py --steps 5 --no-e2e --no-cpu --no-train 2>&1 | tail -2 | cut -c1-2000 | show "aslots$a"; done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.log 2>&1; tail -2 gpurun_out/s8_tests.log
