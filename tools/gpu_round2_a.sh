#!/bin/bash
# Round-2 GPU batch A: parity tests, smoke, benches of BASELINE configs 2-5, graph / split A-B, ncu launch list + full captures.
# Everything is logged under gpurun_out/ (scratch); tools/summarize_profiles.py turns the ncu artefacts into profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
if [ -z "$SKIP_TESTS" ]; then
for f in tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_sampler.py tests/test_gpu_pipeline.py tests/test_gpu_warp.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -rA --maxfail=20 > gpurun_out/${b}_${TAG}.log 2>&1
  echo "== $f exit $?"; grep -E "^\[parity\]|passed|failed|^FAILED|^ERROR|Error:" gpurun_out/${b}_${TAG}.log | tail -${TAIL:-40}
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "== smoke exit $?"; tail -3 gpurun_out/smoke_${TAG}.log
fi
# benches (config 2 = the driver's default line, with the CPU baseline and the per-op profile)
IVID_PROFILE_OPS=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "== bench c2 exit $?"; tail -c 600 gpurun_out/bench_${TAG}_c2.json; tail -3 gpurun_out/bench_${TAG}_c2.err
for c in 3 4 5; do
  IVID_PROFILE_OPS=1 timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err
  echo "== bench c$c exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c$c.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "phase_ms", d["config"]["phase_ms"], "warp", d["roofline"].get("warp"))
except Exception as e:
    print("bench c$c parse failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_c$c.err
done
# A/B: CUDA graph off / output-head split off; small-batch latency with and without the graph
IVID_NO_GRAPH=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_nograph.json 2>/dev/null; echo "== nograph exit $?"
IVID_NO_OUTSPLIT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_nooutsplit.json 2>/dev/null; echo "== nooutsplit exit $?"
for b in 1 2; do
  timeout 600 python bench.py --config 3 --batch $b --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c3_b$b.json 2>/dev/null
  IVID_NO_GRAPH=1 timeout 600 python bench.py --config 3 --batch $b --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c3_b${b}_nograph.json 2>/dev/null
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4f"%d["value"], d["config"].get("phase_ms"))
    except Exception as e: print(f, "parse failed", e)
PY
if [ -z "$SKIP_NCU" ]; then
# ncu: launch list of one short bench run (eager launches: the graph replays the same kernels), then --set full captures
IVID_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 2 -c 2 \
    -o gpurun_out/prof_${TAG}_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_conv.log 2>&1
echo "conv capture exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"gn_apply|attention_kernel|ddpm_step|eps_gather|pack_input" -s 0 -c 12 \
    -o gpurun_out/prof_${TAG}_misc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_misc.log 2>&1
echo "misc capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|resolve_kernel|lanczos|post_|mesh_" -s 60 -c 16 \
    -o gpurun_out/prof_${TAG}_warp -f python tools/bench_warp.py > gpurun_out/ncu_full_${TAG}_warp.log 2>&1
echo "warp capture exit $?"
fi
python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench exit $?"; cat gpurun_out/warp_bench_${TAG}.json
ls -la gpurun_out/*${TAG}* | head -60
