#!/bin/bash
# Round-2 GPU batch AA: 32-bit incremental edge stepping in the rasteriser's small-triangle path: bit-exact tests + timing.
mkdir -p gpurun_out
TAG=${TAG:-r02ab}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest warp+pipeline exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
for i in 1 2; do
  timeout 600 python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}_$i.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/warp_bench_${TAG}_$i.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("aggregate_ms_total","add_view_ms_total","achieved_GBs")})
except Exception as e: print("parse failed", e)
PY
done
timeout 300 python tools/bench_freeview.py > gpurun_out/freeview_bench_${TAG}.json 2>> gpurun_out/warp_bench_${TAG}.err; tail -c 400 gpurun_out/freeview_bench_${TAG}.json
