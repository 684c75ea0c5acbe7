#!/bin/bash
# Round-2 last check of the committed build: whole GPU suite, smoke, headline bench.
mkdir -p gpurun_out
TAG=${TAG:-r02final}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -1 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_c2.json").read().strip().splitlines()[-1]); print("c2: value %.4f e2e %.4f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"], "roofline frac", d["roofline"]["frac"])
PY
