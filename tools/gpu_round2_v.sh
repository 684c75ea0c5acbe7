#!/bin/bash
# Round-2 GPU batch V: compute-sanitizer memcheck over the smoke step, a tiny multiview chain and the warp tests.
mkdir -p gpurun_out
TAG=${TAG:-r02v}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${TAG}_smoke.log 2>&1; echo "== memcheck smoke exit $?"; grep -E "ERROR SUMMARY|Invalid|smoke\]" gpurun_out/sanitizer_${TAG}_smoke.log | tail -6
timeout 1200 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -m pytest tests/test_gpu_warp.py -q -m gpu -x -k "mesh_build or render_matches or forward_backward or postfilter" > gpurun_out/sanitizer_${TAG}_warp.log 2>&1; echo "== memcheck warp exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitizer_${TAG}_warp.log | tail -6
timeout 1200 compute-sanitizer --tool memcheck --print-limit 8 --error-exitcode 9 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sampler.py -q -m gpu -x -k "tiny_unet_vs_reference or backbone_options or ddim_guided or superres" > gpurun_out/sanitizer_${TAG}_unet.log 2>&1; echo "== memcheck unet exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/sanitizer_${TAG}_unet.log | tail -6
