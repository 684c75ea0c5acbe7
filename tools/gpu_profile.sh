#!/bin/bash
# ncu evidence for bench.py (1 GPU): (1) per-launch device times of one short bench run, (2) --set full of the dominant kernels.
mkdir -p gpurun_out
TAG=${TAG:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1200} -c ${COUNT:-800} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list exit $?"
# conv: the 128x128 256->256 residual conv (3rd conv launch of a forward), gn_apply and attention: one launch each
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 2 -c 2 \
    -o gpurun_out/prof_${TAG}_conv -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_conv.log 2>&1
echo "conv capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"gn_apply|attention_kernel|film_table" -s 10 -c 8 \
    -o gpurun_out/prof_${TAG}_misc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_misc.log 2>&1
echo "misc capture exit $?"
ls -la gpurun_out/*.ncu-rep
