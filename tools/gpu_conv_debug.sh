#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
  echo "== IVID_CONV_DEBUG=$d"
  IVID_CONV_DEBUG=$d timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); f=j['roofline']['families']
        print('ms_per_step', round(j['ms_per_step'],2), {k: round(v['ms'],2) for k,v in f.items()})
"
done
