#!/bin/bash
# Round-2 GPU batch G: what the driver runs at round end, on the current build: pytest -m gpu (one process), smoke, default bench,
# reference arm.
mkdir -p gpurun_out
TAG=${TAG:-r02g}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest -m gpu exit $?"; tail -5 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "== bench exit $?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_default.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline"), "roofline frac", d["roofline"]["frac"], "clocks", d["clocks"])
PY
tail -3 gpurun_out/bench_${TAG}_default.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "== reference arm exit $?"; cut -c1-700 gpurun_out/bench_${TAG}_reference.json; tail -3 gpurun_out/bench_${TAG}_reference.err
