#!/bin/bash
# Round-2 GPU batch J: rasteriser occupancy A/B (smem-resident big triangle, 96 regs vs 122), embedding taps test,
# two-stream split-batch overlap probe, clock sampler check.
mkdir -p gpurun_out
TAG=${TAG:-r02j}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_unet.py tests/test_gpu_sampler.py -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
grep "\[tap\].*emb" gpurun_out/pytest_gpu_${TAG}.log | head -3
for v in "" "IVID_RASTER_REGTRI=1" "" "IVID_RASTER_REGTRI=1"; do
  env $v timeout 600 python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}_${v:-smemtri}.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench ${v:-smemtri} exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/warp_bench_${TAG}_${v:-smemtri}.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("aggregate_ms_total","add_view_ms_total","achieved_GBs")})
except Exception as e: print("parse failed", e)
PY
done
for a in "32 2" "32 4"; do
  timeout 600 python tools/micro/overlap_probe.py $a 2> gpurun_out/overlap_${TAG}.err | tee -a gpurun_out/overlap_${TAG}.json; echo "== overlap $a exit $?"
done
tail -3 gpurun_out/overlap_${TAG}.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2.json").read().strip().splitlines()[-1]); print("c2 ms/step %.3f"%d["ms_per_step"], "clocks", d["clocks"], "e2e", d["e2e"]["value"])
except Exception as e: print("parse failed", e)
PY
