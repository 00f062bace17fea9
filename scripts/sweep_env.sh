#!/bin/bash
# usage: sweep_env.sh <workload> "ENV1=a ENV2=b" "ENV1=c" ...   - one bench line per environment, stage times only
wl=$1; shift
for e in "$@"; do
  env VQB_EXPERIMENTS=1 $e timeout 200 python bench.py --workload $wl --no-e2e --no-cpu --no-train --no-sampler 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('%-44s' % '$e', 'search %.4f tail %.4f prep %.4f step %.4f' % (s['search'], s['tail'], s['prep'], d['ms_per_step']), d['shortlist']['fallback_frames_per_step'])"
done
