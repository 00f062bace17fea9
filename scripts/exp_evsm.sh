#!/bin/bash
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items()})
"; }
for e in 0 1 2 3; do VQB_TC_EVSM=$e timeout 120 python bench.py --steps 50 --no-e2e --no-cpu --no-train --workload cfg1 2>/dev/null | show "cfg1 evsm$e"; done
VQB_TC_EVSM=0 VQB_TC_STAGES=4 timeout 120 python bench.py --steps 50 --no-e2e --no-cpu --no-train --workload cfg1 2>/dev/null | show "cfg1 evsm0 st4"
VQB_TC_EVSM=3 timeout 120 python bench.py --steps 50 --no-e2e --no-cpu --no-train --no-sampler --workload cfg1 2>/dev/null | show "cfg1 evsm3 nosampler"
VQB_TC_EVSM=0 timeout 120 python bench.py --steps 50 --no-e2e --no-cpu --no-train --no-sampler --workload cfg1 2>/dev/null | show "cfg1 evsm0 nosampler"
