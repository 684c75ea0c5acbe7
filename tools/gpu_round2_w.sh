#!/bin/bash
# Round-2 GPU batch Y: GroupNorm fold with four transform warps (320-thread tap-reuse kernel): parity + A/B.
mkdir -p gpurun_out
TAG=${TAG:-r02y}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_ops.py -q -m gpu -x -k "fold or slab or (conv_matches and slab)" > gpurun_out/pytest_gpu_${TAG}_fold.log 2>&1; echo "== pytest fold/slab exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}_fold.log
IVID_FOLD=1 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -k "layerwise or three_timesteps" > gpurun_out/pytest_gpu_${TAG}_fold2.log 2>&1; echo "== pytest IVID_FOLD=1 exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}_fold2.log
for v in "IVID_FOLD=0" "IVID_FOLD=1" "IVID_FOLD=0" "IVID_FOLD=1"; do
  env $v timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_${TAG}_c2_${v}.json" 2>gpurun_out/bench_${TAG}.err
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
