#!/bin/bash
export VQB_EXPERIMENTS=1
for pass in 1 2; do
for e in "VQB_TC_SLEEP=0" "VQB_TC_SLEEP=40" "VQB_TC_SLEEP=150"; do
  env $e timeout 200 python bench.py --no-e2e --no-cpu --no-train --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('%-20s' % '$e', 'search %.3f tail %.3f step %.3f' % (s['search'], s['tail'], d['ms_per_step']), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
done; done
