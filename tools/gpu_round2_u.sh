#!/bin/bash
# Round-2 GPU batch U (FINAL build): full parity suite, smoke, default bench (with CPU baseline), configs 3-5, reference arm,
# full-pipeline runs, launch list + ncu captures of the conv and attention kernels.
mkdir -p gpurun_out
TAG=${TAG:-r02u}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
IVID_PROFILE_OPS=1 timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench c2 exit $?"
cp gpurun_out/per_op_profile_c2.json gpurun_out/per_op_${TAG}_c2.json 2>/dev/null
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "== bench c$c exit $?"
done
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "== reference arm exit $?"
for c in 2 5; do
  timeout 1200 python bench.py --config $c --full --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}_full.json 2> gpurun_out/bench_${TAG}_c${c}_full.err; echo "== full c$c exit $?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms/step %.3f"%d["ms_per_step"], "value %.5f"%d["value"], "e2e %.5f"%d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
    except Exception as e: print(f, "parse failed", e)
PY
IVID_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_${TAG}_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_conv.log 2>&1
echo "conv capture exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 0 -c 6 \
    -o gpurun_out/prof_${TAG}_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_attn.log 2>&1
echo "attention capture exit $?"
ls -la gpurun_out/*${TAG}* | head -40
