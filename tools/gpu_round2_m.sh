#!/bin/bash
# Round-2 GPU batch N: attention v2 after the barrier fix: determinism (kernel alone, whole forward), parity tests, timing.
mkdir -p gpurun_out
TAG=${TAG:-r02n}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 600 python tools/micro/attn_determinism.py 6000 2>> gpurun_out/attn_det_${TAG}.err | tee -a gpurun_out/attn_det_${TAG}.json; echo "== attn determinism exit $?"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" > gpurun_out/pytest_gpu_${TAG}_attn.log 2>&1; echo "== pytest attention exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}_attn.log
timeout 600 python tools/micro/determinism_probe.py 32 200 0 2>> gpurun_out/attn_det_${TAG}.err | tee gpurun_out/det_${TAG}_graph_many.json | cut -c1-300; echo "== forward determinism exit $?"
IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2.json 2>gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3), round(v["tflops"])) for k,v in f.items() if k.startswith("att")}, d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/attn_det_${TAG}.err
