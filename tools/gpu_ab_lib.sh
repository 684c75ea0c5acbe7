#!/bin/bash
# A/B of two builds of the library on the same box: tools/gpu_ab_lib.sh other.so   (the in-tree build is "new")
OTHER=$1
cp ivid_b200/libivid_b200.so /tmp/new.so
for rep in 1 2; do for v in new other; do
  if [ $v = new ]; then cp /tmp/new.so ivid_b200/libivid_b200.so; else cp $OTHER ivid_b200/libivid_b200.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); f=j['roofline']['families']
        print('$v', 'ms_per_step', round(j['ms_per_step'],2), 'clk', j['clocks']['sm_mhz'], {k: round(x['ms'],2) for k,x in f.items() if x['ms']>0.3})
"
done; done
cp /tmp/new.so ivid_b200/libivid_b200.so
