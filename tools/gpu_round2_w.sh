#!/bin/bash
# Round-2 GPU batch W: probe of unaligned (base-offset) UMMA descriptor starts through slab mode 2 / 3.
mkdir -p gpurun_out
TAG=${TAG:-r02w}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 300 python tools/micro/slab_probe.py 2> gpurun_out/slab_probe_${TAG}.err | tee gpurun_out/slab_probe_${TAG}.json; echo "== probe exit $?"; tail -3 gpurun_out/slab_probe_${TAG}.err
