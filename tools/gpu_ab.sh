#!/bin/bash
# A/B of an environment knob under the same box / power state: tools/gpu_ab.sh VAR v1 v2 ...
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); f=j['roofline']['families']
        print('$VAR=$v', 'ms_per_step', round(j['ms_per_step'],2), 'clk', j['clocks']['sm_mhz'], {k: round(x['ms'],2) for k,x in f.items() if x['ms']>0.5})
"
done; done
