#!/bin/bash
# Round-2 GPU batch Z (final): whole GPU suite, smoke, benches of the four configurations on the final build, fold diagnostic.
mkdir -p gpurun_out
TAG=${TAG:-r02z}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
timeout 600 python tools/micro/fold_diff.py 2> gpurun_out/fold_diff_${TAG}.err > gpurun_out/fold_diff_${TAG}.json; echo "== fold diff exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/fold_diff_${TAG}.json")); print("fold vs apply eps", d["eps_fold_vs_apply"], "fold vs fold", d["eps_fold_vs_fold"])
    for b in d["blocks"][:14]: print(b)
except Exception as e: print("parse failed", e)
PY
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "== bench c$c exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c${c}.json").read().strip().splitlines()[-1]); print("c$c: value %.4f e2e %.4f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
except Exception as e: print("parse failed", e)
PY
done
