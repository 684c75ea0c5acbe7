#!/bin/bash
# Round-2 GPU batch C: attention v2 (double-buffered S) + warp shade/resolve split + coefficient-form edges: parity, A/B, ncu.
mkdir -p gpurun_out
TAG=${TAG:-r02c}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for f in tests/test_gpu_ops.py tests/test_gpu_warp.py tests/test_gpu_unet.py tests/test_gpu_sampler.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -rA --maxfail=20 > gpurun_out/${b}_${TAG}.log 2>&1
  echo "== $f exit $?"; grep -E "attention|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/${b}_${TAG}.log | tail -25
done
python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench exit $?"; cat gpurun_out/warp_bench_${TAG}.json; tail -3 gpurun_out/warp_bench_${TAG}.err
for v in "" "IVID_ATTN_V1=1"; do
  for c in 2 5; do
    env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}_${v:-v2}.json 2>gpurun_out/bench_${TAG}_c${c}_${v:-v2}.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c${c}_${v:-v2}.json").read().strip().splitlines()[-1])
    print("c$c ${v:-v2}: ms/step %.3f"%d["ms_per_step"], "attention", d["roofline"]["families"].get("attention"))
except Exception as e:
    print("parse failed", e)
PY
  done
done
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 0 -c 6 \
    -o gpurun_out/prof_${TAG}_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_attn.log 2>&1
echo "attention capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|shade_kernel|resolve_kernel" -s 18 -c 6 \
    -o gpurun_out/prof_${TAG}_warp -f python tools/bench_warp.py > gpurun_out/ncu_full_${TAG}_warp.log 2>&1
echo "warp capture exit $?"
ls -la gpurun_out/*${TAG}* | head -40
