#!/bin/bash
for f in "$@"; do python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(sys.argv[1].split('/')[-1],'step',round(d['ms_per_step'],3),'tc', round(r['kernel_ms'],3), 'mhz', d['clocks']['sm_mhz'], 'train', (d.get('train_step') or {}).get('ms_per_step'), 'emb', d['losses']['embedding'])
    elif 'Error' in l or 'error' in l: print(l.strip()[:200])
PY
done
