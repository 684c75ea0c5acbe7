#!/bin/bash
# Round-2 GPU batch I: lean SiLU A/B, staged pack_input, full runs with one warm-up batch.
mkdir -p gpurun_out
TAG=${TAG:-r02i}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_sampler.py -q -m gpu -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
for v in "" "IVID_SILU_WRAPPED=1" "" "IVID_SILU_WRAPPED=1"; do
  env $v IVID_PROFILE_OPS=1 timeout 600 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_${v:-lean}.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c2_${v:-lean}.json").read().strip().splitlines()[-1])
    f=d["roofline"]["families"]
    print("c2 ${v:-lean}: ms/step %.3f"%d["ms_per_step"], {k:(v["launches"], round(v["ms"],3)) for k,v in f.items() if k.startswith("gn") or k.startswith("pack")})
except Exception as e:
    print("parse failed", e)
PY
done
for c in 2 5; do
  timeout 1200 python bench.py --config $c --full --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}_full.json 2> gpurun_out/bench_${TAG}_c${c}_full.err; echo "== full c$c exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c${c}_full.json").read().strip().splitlines()[-1]); print("full c$c: s/batch %.3f"%(d["ms_per_step"]/1e3), "samples/s %.4f"%d["value"])
except Exception as e: print("parse failed", e)
PY
done
