#!/bin/bash
# Round-2 GPU batch H (final build): full parity suite, smoke, all four configurations, full-pipeline runs, launch list + ncu captures.
mkdir -p gpurun_out
TAG=${TAG:-r02h}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1500 python -m pytest tests/ -q -m gpu -s -rA > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "== pytest -m gpu exit $?"; grep -E "fused|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu_${TAG}.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke_${TAG}.log
IVID_PROFILE_OPS=1 timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench c2 exit $?"
for c in 3 4 5; do
  IVID_PROFILE_OPS=1 timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "== bench c$c exit $?"
done
IVID_NO_FUSED_STEP=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_nofusedstep.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "== reference arm exit $?"
for c in 2 3 5; do
  timeout 1200 python bench.py --config $c --full --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_${TAG}_c${c}_full.json 2> gpurun_out/bench_${TAG}_c${c}_full.err; echo "== full c$c exit $?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms/step %.3f"%d["ms_per_step"], "value %.5f"%d["value"], "e2e %.5f"%d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"), d["config"].get("phase_ms"))
    except Exception as e: print(f, "parse failed", e)
PY
IVID_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_${TAG}_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_conv.log 2>&1
echo "conv capture exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"head_step_kernel|eps_gather|pack_input|conv_gemm_kernel<64" -s 0 -c 4 \
    -o gpurun_out/prof_${TAG}_head -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_head.log 2>&1
echo "head capture exit $?"
ls -la gpurun_out/*${TAG}* | head -40
