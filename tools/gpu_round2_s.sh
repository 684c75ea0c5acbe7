#!/bin/bash
# Round-2 GPU batch S: hybrid conv schedule (contiguous ranges on single-column-block layers) A/B, backbone option parity.
mkdir -p gpurun_out
TAG=${TAG:-r02s}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -k "backbone_options or layerwise or real_config" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
for v in "IVID_X=0" "IVID_CONV_STRIDED=1" "IVID_X=0" "IVID_CONV_STRIDED=1"; do
  env $v timeout 600 python bench.py --config 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v}.json 2>gpurun_out/bench_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v}.json").read().strip().splitlines()[-1])
    print("c2 ${v}: ms/step %.3f value %.4f"%(d["ms_per_step"], d["value"]), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
tail -3 gpurun_out/bench_${TAG}.err
