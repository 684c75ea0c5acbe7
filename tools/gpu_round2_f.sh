#!/bin/bash
# Round-2 GPU batch F: GroupNorm SiLU with the FMA-pipe reciprocal (A/B vs SFU), N-tile choice on the low-resolution levels.
mkdir -p gpurun_out
TAG=${TAG:-r02f}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -s -rA --maxfail=30 > gpurun_out/test_gpu_ops_${TAG}.log 2>&1
echo "== ops exit $?"; grep -E "group_norm|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/test_gpu_ops_${TAG}.log | tail -16
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s -rA --maxfail=30 -k "golden or real_config" > gpurun_out/test_gpu_unet_${TAG}.log 2>&1
echo "== unet exit $?"; grep -E "eps rel|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/test_gpu_unet_${TAG}.log | tail -12
for v in "" "IVID_SILU_SFU=1" "IVID_BN128_COST=2.0" "" "IVID_SILU_SFU=1" "IVID_BN128_COST=2.0"; do
  env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v:-default}.json 2>gpurun_out/bench_${TAG}_c2_${v:-default}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v:-default}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2 ${v:-default}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("conv") or k.startswith("gn")})
except Exception as e:
    print("parse failed", e)
PY
  tail -2 gpurun_out/bench_${TAG}_c2_${v:-default}.err
  cp gpurun_out/per_op_profile_c2.json "gpurun_out/per_op_${TAG}_${v:-default}.json" 2>/dev/null
done
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"gn_apply" -s 1 -c 3 \
    -o gpurun_out/prof_${TAG}_gn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_gn.log 2>&1
echo "gn capture exit $?"
ls -la gpurun_out/*${TAG}* | head -20
