#!/bin/bash
# Full GPU check used during development: parity tests (one process per file so a sticky CUDA error cannot poison the
# next file), smoke, a short bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for f in ${TEST_FILES:-tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_sampler.py}; do
  b=$(basename $f .py)
  timeout ${TEST_TIMEOUT:-600} python -m pytest $f -q -m gpu -s ${PYTEST_ARGS} > gpurun_out/$b.log 2>&1
  echo "== $f exit $?"; grep -E "^\[parity\]|passed|failed|Error|error" gpurun_out/$b.log | tail -${TAIL:-60}
done
if [ -z "$SKIP_SMOKE" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -3 gpurun_out/smoke.log
fi
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
