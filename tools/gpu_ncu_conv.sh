#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s ${KSKIP:-2} -c ${KCOUNT:-1} \
    -o gpurun_out/prof_${TAG:-conv} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG:-conv}.log 2>&1
echo "exit $?"; ls -la gpurun_out/*.ncu-rep
