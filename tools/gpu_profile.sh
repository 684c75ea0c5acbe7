#!/bin/bash
# ncu evidence for bench.py (1 GPU): (1) per-launch device times of one short bench run, (2) --set full of the dominant kernels.
mkdir -p gpurun_out
TAG=${TAG:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-800} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-conv_gemm_kernel} -s ${KSKIP:-300} -c ${KCOUNT:-3} \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
