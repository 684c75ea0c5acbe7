#!/bin/bash
# Round-2 GPU batch L: run-to-run determinism of the forward at batch 32 (eager vs CUDA graph), localised by taps.
mkdir -p gpurun_out
TAG=${TAG:-r02l}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
IVID_NO_GRAPH=1 timeout 900 python tools/micro/determinism_probe.py 32 3 1 2> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_eager.json; echo "== eager exit $?"
timeout 900 python tools/micro/determinism_probe.py 32 4 1 2>> gpurun_out/det_${TAG}.err | tee gpurun_out/det_${TAG}_graph.json; echo "== graph exit $?"
tail -3 gpurun_out/det_${TAG}.err
