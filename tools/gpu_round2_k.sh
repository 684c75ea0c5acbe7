#!/bin/bash
# Round-2 GPU batch K: 3x3 tap-reuse ("slab") conv kernel: op parity, UNet parity with IVID_SLAB=1, step time A/B; batch invariance probe.
mkdir -p gpurun_out
TAG=${TAG:-r02k}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 300 python tools/micro/batch_invariance.py 32 L 2> gpurun_out/binv_${TAG}.err | tee gpurun_out/binv_${TAG}.json; echo "== binv exit $?"; tail -2 gpurun_out/binv_${TAG}.err
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv" > gpurun_out/pytest_gpu_${TAG}_ops.log 2>&1; echo "== pytest ops exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}_ops.log
grep -i "slab" gpurun_out/pytest_gpu_${TAG}_ops.log | head -12
IVID_SLAB=1 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}_unet_slab.log 2>&1; echo "== pytest unet slab exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}_unet_slab.log
for v in "IVID_SLAB=0" "IVID_SLAB=1" "IVID_SLAB=0" "IVID_SLAB=1"; do
  env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v}.json 2>gpurun_out/bench_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2 ${v}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("conv") or k.startswith("gn")}, d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
tail -3 gpurun_out/bench_${TAG}.err
