#!/bin/bash
# Round-2 GPU batch Q: backbone option parity (noshift / plainconv / plainpool), conv epilogue attribution (IVID_CONV_DEBUG bits).
mkdir -p gpurun_out
TAG=${TAG:-r02q}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -s -k "backbone_options" > gpurun_out/pytest_gpu_${TAG}_options.log 2>&1; echo "== pytest options exit $?"; grep "^\[\|passed\|failed\|Error\|rel " gpurun_out/pytest_gpu_${TAG}_options.log | tail -40
for d in 0 1 2 9; do
  IVID_CONV_DEBUG=$d IVID_PROFILE_OPS=1 timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_dbg$d.json 2>gpurun_out/bench_${TAG}.err
  cp gpurun_out/per_op_profile_c2.json gpurun_out/per_op_${TAG}_dbg$d.json 2>/dev/null
  python - <<PY
import json, collections
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_dbg$d.json").read().strip().splitlines()[-1])
    ops=json.load(open("gpurun_out/per_op_${TAG}_dbg$d.json"))
    agg=collections.defaultdict(float)
    for fam,desc,ms,fl,by in ops:
        if fam.startswith("conv_gemm<256>") and "128x128" in desc: agg[desc]+=ms
    print("debug $d: ms/step %.3f"%d["ms_per_step"], {k: round(v,3) for k,v in sorted(agg.items())})
except Exception as e:
    print("parse failed", e)
PY
done
tail -3 gpurun_out/bench_${TAG}.err
