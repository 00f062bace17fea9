#!/bin/bash
# timing-only experiments on the small-codebook shape: which part of tc_search_kernel's epilogue costs what
# (VQB_TAIL_DBG bits 256 / 512 / 1024 / 2048 switch parts of it off; results are then wrong, only the stage time is read)
wl=${1:-cfg2}
for spec in "0" "256" "512" "1024" "2048" "768" "3584" ; do
  VQB_TIMING_EXPERIMENTS=1 VQB_TAIL_DBG=$spec timeout 120 python bench.py --workload $wl --no-e2e --no-cpu --no-train --no-sampler 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dbg=$spec', 'search %.4f ms' % d['stage_ms_per_step']['search'], 'step %.4f' % d['ms_per_step'], d['shortlist'])"
done
for e in "VQB_TC_FUSE=0" "VQB_TC_MODE=1" "VQB_TC_EVSM=0" "VQB_TC_MODE=1 VQB_TC_CLUSTER=1"; do
  env $e timeout 120 python bench.py --workload $wl --no-e2e --no-cpu --no-train --no-sampler 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$e', 'search %.4f ms' % d['stage_ms_per_step']['search'], 'prep %.4f' % d['stage_ms_per_step']['prep'], 'step %.4f' % d['ms_per_step'])"
done
