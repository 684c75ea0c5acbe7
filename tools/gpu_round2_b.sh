#!/bin/bash
# Round-2 GPU batch B: warp parity (new rasteriser inner loop, SimpleRenderer / forward_backward_warp), warp bench, ncu of the
# warp kernels and of attention.
mkdir -p gpurun_out
TAG=${TAG:-r02b}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for f in tests/test_gpu_warp.py tests/test_gpu_pipeline.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -rA --maxfail=20 > gpurun_out/${b}_${TAG}.log 2>&1
  echo "== $f exit $?"; grep -E "^\[parity\]|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/${b}_${TAG}.log | tail -40
done
python tools/bench_warp.py > gpurun_out/warp_bench_${TAG}.json 2> gpurun_out/warp_bench_${TAG}.err; echo "== warp bench exit $?"; cat gpurun_out/warp_bench_${TAG}.json; tail -3 gpurun_out/warp_bench_${TAG}.err
ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|resolve_kernel" -s 12 -c 4 \
    -o gpurun_out/prof_${TAG}_warp -f python tools/bench_warp.py > gpurun_out/ncu_full_${TAG}_warp.log 2>&1
echo "warp capture exit $?"
IVID_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 0 -c 6 \
    -o gpurun_out/prof_${TAG}_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_attn.log 2>&1
echo "attention capture exit $?"
ls -la gpurun_out/*${TAG}*
