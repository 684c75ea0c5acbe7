// Timestep / class embedding path of the ADM UNet (fp32, CUDA cores; negligible FLOPs, weight-bandwidth bound).
//   reference: PosEncoding adm.py:28-33, time_embed adm.py:357-362, label_emb adm.py:365,547-555,
//              ResBlock emb_layers (SiLU -> Linear) adm.py:174-177,211
// All 35 emb_layers Linears are evaluated as ONE [N,E] x [E, sum(2*Cout)] product per forward (the "FiLM table").
#pragma once
#include "common.cuh"

namespace ivid {

// emb0[n][:] = [cos(t*f) | sin(t*f)]
__global__ void posenc_kernel(const int64_t* __restrict__ t, int Nt, const float* __restrict__ freqs, int half,
                              float* __restrict__ out, int N) {
  const int n = blockIdx.x;
  const float tv = static_cast<float>(t[n % Nt]);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = tv * freqs[i];
    out[static_cast<size_t>(n) * 2 * half + i] = cosf(a);
    out[static_cast<size_t>(n) * 2 * half + half + i] = sinf(a);
  }
}

// out[n][o] = bias[o] + sum_k act(in[n][k]) * W[o][k]   (+ class embedding rows when label_emb != null)
// One warp per output feature, up to 32 batch rows per warp pass; grid = (ceil(O/8), ceil(N/32)), block = 256.
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out, int N,
                                                          int K, int O, int silu_in,
                                                          const float* __restrict__ label_emb,
                                                          const int64_t* __restrict__ classes, int Ncls) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  const int n0 = blockIdx.y * 32;
  if (o >= O) return;
  const int rows = min(32, N - n0);
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = __ldg(W + static_cast<size_t>(o) * K + k);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      if (r < rows) {
        float v = __ldg(in + static_cast<size_t>(n0 + r) * K + k);
        if (silu_in) v = v / (1.0f + expf(-v));
        acc[r] = fmaf(v, w, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    float v = acc[r];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[r] = v;
  }
  if (lane == 0) {
    for (int r = 0; r < rows; ++r) {
      float v = acc[r] + bias[o];
      if (label_emb != nullptr && classes != nullptr) {
        const int64_t c = classes[(n0 + r) % Ncls];
        if (c >= 0) v += label_emb[static_cast<size_t>(c) * O + o];   // null class (-1) contributes zero (adm.py:551-553)
      }
      out[static_cast<size_t>(n0 + r) * O + o] = v;
    }
  }
}

}  // namespace ivid
