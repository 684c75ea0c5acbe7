#!/bin/bash
# Round-2 GPU batch T: residual-prefetch cursor (no per-chunk integer divisions in the fp32 epilogue): parity + timing.
mkdir -p gpurun_out
TAG=${TAG:-r02t}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_sampler.py -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
for v in "IVID_X=0" "IVID_X=1"; do
  env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v}.json 2>gpurun_out/bench_${TAG}.err
  cp gpurun_out/per_op_profile_c2.json gpurun_out/per_op_${TAG}_${v}.json 2>/dev/null
  python - <<PY
import json, collections
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    ops=json.load(open("gpurun_out/per_op_${TAG}_${v}.json"))
    res=sum(ms for fam,desc,ms,fl,by in ops if fam.startswith("conv") and " res" in desc)
    print("c2 ${v}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("conv") or k.startswith("gn")}, "res convs %.3f ms"%res, d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
timeout 600 python bench.py --config 2 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph run: ms/step %.3f value %.4f'%(d['ms_per_step'], d['value']), d['clocks'])"
tail -3 gpurun_out/bench_${TAG}.err
