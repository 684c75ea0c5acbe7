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
// grid = (ceil(O/64), ceil(N/32)), block = 256.
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out, int N,
                                                          int K, int O, int silu_in,
                                                          const float* __restrict__ label_emb,
                                                          const int64_t* __restrict__ classes, int Ncls) {
  // Block = 64 output features x 32 batch rows; K walked in 64-wide chunks staged in shared memory (inputs and weights
  // both read from HBM/L2 exactly once per block).  Lane = batch row, warp = 8 output features.
  constexpr int KC = 64;
  __shared__ float s_in[32][KC + 1];
  __shared__ float s_w[64][KC + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * 64;
  const int n0 = blockIdx.y * 32;
  const int rows = min(32, N - n0);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int kc = 0; kc < K; kc += KC) {
    for (int idx = threadIdx.x; idx < 32 * KC; idx += 256) {
      const int r = idx / KC, c = idx % KC;
      float v = 0.f;
      if (r < rows && kc + c < K) {
        v = __ldg(in + static_cast<size_t>(n0 + r) * K + kc + c);
        if (silu_in) v = v / (1.0f + expf(-v));
      }
      s_in[r][c] = v;
    }
    for (int idx = threadIdx.x; idx < 64 * KC; idx += 256) {
      const int r = idx / KC, c = idx % KC;
      s_w[r][c] = (o0 + r < O && kc + c < K) ? __ldg(W + static_cast<size_t>(o0 + r) * K + kc + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      const float x = s_in[lane][k];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(x, s_w[warp * 8 + i][k], acc[i]);
    }
    __syncthreads();
  }
  if (lane < rows) {
    const int n = n0 + lane;
    int64_t cls = -1;
    if (label_emb != nullptr && classes != nullptr) cls = classes[n % Ncls];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int o = o0 + warp * 8 + i;
      if (o < O) {
        float v = acc[i] + bias[o];
        if (cls >= 0) v += label_emb[static_cast<size_t>(cls) * O + o];   // null class (-1) contributes zero (adm.py:551-553)
        out[static_cast<size_t>(n) * O + o] = v;
      }
    }
  }
}

}  // namespace ivid
