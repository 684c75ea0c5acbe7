#!/bin/bash
# Round-2 GPU batch O: full validation of the build after the attention fix: whole GPU test-suite, smoke, determinism of the
# conditional / SR forwards, default bench + configs 3-5.
mkdir -p gpurun_out
TAG=${TAG:-r02o}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest exit $?"; tail -5 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -3 gpurun_out/smoke_${TAG}.log
timeout 600 python tools/micro/determinism_probe.py 32 150 0 Lc 2>> gpurun_out/det_${TAG}.err | cut -c1-200 | tee gpurun_out/det_${TAG}_Lc.json; echo "== Lc determinism exit $?"
timeout 600 python tools/micro/determinism_probe.py 16 100 0 SR 2>> gpurun_out/det_${TAG}.err | cut -c1-200 | tee gpurun_out/det_${TAG}_SR.json; echo "== SR determinism exit $?"
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench default exit $?"; cat gpurun_out/bench_${TAG}_c2.json | cut -c1-600
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}.json 2> gpurun_out/bench_${TAG}_c${c}.err; echo "== bench c$c exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c${c}.json").read().strip().splitlines()[-1]); print("c$c: value %.4f e2e %.4f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"], d["config"].get("phases_ms"))
except Exception as e: print("parse failed", e)
PY
done
tail -3 gpurun_out/det_${TAG}.err
