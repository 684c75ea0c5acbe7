#!/bin/bash
# Round-2 GPU batch E: cluster-multicast conv (low-resolution levels): parity + same-box A/B.
mkdir -p gpurun_out
TAG=${TAG:-r02e}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -s -rA --maxfail=30 > gpurun_out/test_gpu_ops_${TAG}.log 2>&1
echo "== ops exit $?"; grep -E "multicast|8x8|16x16|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/test_gpu_ops_${TAG}.log | tail -40
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s -rA --maxfail=30 -k "tiny or real_config" > gpurun_out/test_gpu_unet_${TAG}.log 2>&1
echo "== unet exit $?"; grep -E "eps rel|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/test_gpu_unet_${TAG}.log | tail -30
for v in "" "IVID_NO_MC=1" "" "IVID_NO_MC=1"; do
  env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v:-mc}.json 2>gpurun_out/bench_${TAG}_c2_${v:-mc}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v:-mc}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2 ${v:-mc}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("conv")})
except Exception as e:
    print("parse failed", e)
PY
  tail -2 gpurun_out/bench_${TAG}_c2_${v:-mc}.err
  cp gpurun_out/per_op_profile_c2.json gpurun_out/per_op_${TAG}_${v:-mc}.json 2>/dev/null
done
ls -la gpurun_out/*${TAG}* | head
