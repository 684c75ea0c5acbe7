#!/bin/bash
# Round-2 GPU batch D: warp (stash, add_view timing, device-side free-view resolve), config 4 bench, full-pipeline runs.
mkdir -p gpurun_out
TAG=${TAG:-r02d}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for f in tests/test_gpu_warp.py tests/test_gpu_ops.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -rA --maxfail=20 > gpurun_out/${b}_${TAG}.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/${b}_${TAG}.log | tail -15
done
python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench exit $?"; cat gpurun_out/warp_bench_${TAG}.json; tail -3 gpurun_out/warp_bench_${TAG}.err
python tools/bench_freeview.py > gpurun_out/freeview_bench_${TAG}.json 2> gpurun_out/freeview_bench_${TAG}.err; echo "== freeview exit $?"; cat gpurun_out/freeview_bench_${TAG}.json; tail -3 gpurun_out/freeview_bench_${TAG}.err
timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err; echo "== c4 exit $?"; tail -c 900 gpurun_out/bench_${TAG}_c4.json | head -c 900; tail -3 gpurun_out/bench_${TAG}_c4.err
if [ -n "$FULL" ]; then
for c in 2 3 5; do
  timeout 1200 python bench.py --config $c --full --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}_full.json 2> gpurun_out/bench_${TAG}_c${c}_full.err
  echo "== full c$c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c${c}_full.json").read().strip().splitlines()[-1]); print("full c$c: s/batch %.2f"%(d["ms_per_step"]/1e3), "samples/s %.4f"%d["value"])
except Exception as e: print("parse failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_c${c}_full.err
done
fi
ls -la gpurun_out/*${TAG}* | head -30
