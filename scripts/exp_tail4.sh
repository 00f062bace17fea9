#!/bin/bash
for flags in "" "--no-resid" "--no-q" "--no-resid --no-q"; do
  VQB_TAIL_VARIANT=${V:-1} timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train $flags 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$flags', 'step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['clocks']['sm_mhz'])
"
done
