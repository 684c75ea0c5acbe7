#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 300 python tools/micro/latency_small_batch.py 2> gpurun_out/latency_r02.err | tee gpurun_out/latency_r02_graph.json
IVID_NO_GRAPH=1 timeout 300 python tools/micro/latency_small_batch.py 2>> gpurun_out/latency_r02.err | tee gpurun_out/latency_r02_nograph.json
tail -3 gpurun_out/latency_r02.err
