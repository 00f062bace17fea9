#!/bin/bash
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['clocks'])
"; }
VQB_TAIL_VARIANT=1 timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train 2>/dev/null | show sampler
VQB_TAIL_VARIANT=1 timeout 120 python bench.py --steps 4 --no-e2e --no-cpu --no-train --no-sampler 2>/dev/null | show nosampler
VQB_TAIL_VARIANT=1 python scripts/bench_tail_only.py 8192 4 2>&1 | tail -1 | cut -c1-160
