#!/bin/bash
# Round-2 GPU batch W: unaligned (base-offset) UMMA descriptor starts (slab mode 2 / 3) and the GroupNorm fold built on them.
mkdir -p gpurun_out
TAG=${TAG:-r02x}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 300 python tools/micro/slab_probe.py 2> gpurun_out/slab_probe_${TAG}.err | tee gpurun_out/slab_probe_${TAG}.json; echo "== probe exit $?"; tail -3 gpurun_out/slab_probe_${TAG}.err
IVID_FOLD=1 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -s -k "layerwise or real_config or three_timesteps" > gpurun_out/pytest_gpu_${TAG}_fold.log 2>&1; echo "== pytest fold exit $?"; grep -E "^\[tap\] large|passed|failed|Error|rel " gpurun_out/pytest_gpu_${TAG}_fold.log | tail -30
for v in "IVID_FOLD=0" "IVID_FOLD=1" "IVID_FOLD=1 IVID_SLAB=2" "IVID_FOLD=0" "IVID_FOLD=1"; do
  env $v timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_${TAG}_c2_${v// /_}.json" 2>gpurun_out/bench_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v// /_}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2 ${v}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("conv") or k.startswith("gn")}, d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
tail -3 gpurun_out/bench_${TAG}.err
