#!/bin/bash
# Round-2 GPU batch L3: localise the rare run-to-run difference of the batch-32 forward with per-block taps (many runs).
mkdir -p gpurun_out
TAG=${TAG:-r02l3}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 1200 python tools/micro/determinism_probe.py 32 80 1 2> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_graph.json; echo "== graph+taps exit $?"
timeout 600 python tools/micro/determinism_probe.py 32 120 0 2>> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_graph_many.json; echo "== graph many exit $?"
IVID_ATTN_V1=1 timeout 600 python tools/micro/determinism_probe.py 32 120 0 2>> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_attnv1.json; echo "== attn v1 exit $?"
timeout 600 python tools/micro/determinism_probe.py 8 150 0 2>> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_n8.json; echo "== N=8 exit $?"
tail -3 gpurun_out/det_${TAG}.err
