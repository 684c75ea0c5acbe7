#!/bin/bash
# builds the stand-alone micro-benchmarks (binaries are git-ignored but travel to the GPU box with gpurun)
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bench_gn bench_gn.cu
