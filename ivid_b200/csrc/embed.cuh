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

// Register-tiled variant for K % 32 == 0 (every real configuration): block = 32 batch rows x 32*OPT output features,
// thread = 4 rows x OPT features (feature = tx + 32*i, so weight reads are conflict-free and output stores coalesced).
// The next K chunk is prefetched into registers while the current one is consumed; weights are read from HBM once.
template <int OPT>
__global__ void __launch_bounds__(256) linear_tiled_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ out, int N,
                                                           int K, int O, int silu_in,
                                                           const float* __restrict__ label_emb,
                                                           const int64_t* __restrict__ classes, int Ncls) {
  constexpr int KC = 32, TO = 32 * OPT, LDW = KC + 4;
  __shared__ __align__(16) float s_w[TO * LDW];
  __shared__ __align__(16) float s_in[KC * 32];        // [k][row]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o0 = blockIdx.x * TO, n0 = blockIdx.y * 32;
  const int rows = min(32, N - n0);
  float acc[4][OPT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < OPT; ++i) acc[j][i] = 0.f;
  float4 wreg[OPT];
  float xreg[4];
  auto fetch = [&](int kc) {
#pragma unroll
    for (int u = 0; u < OPT; ++u) {
      const int idx = threadIdx.x + u * 256, r = idx >> 3, c4 = idx & 7;
      wreg[u] = (o0 + r < O) ? ldg_f4(W + static_cast<size_t>(o0 + r) * K + kc + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = threadIdx.x + u * 256, r = idx >> 5, c = idx & 31;       // consecutive lanes read consecutive k
      float v = 0.f;
      if (r < rows) {
        v = __ldg(in + static_cast<size_t>(n0 + r) * K + kc + c);
        if (silu_in) v = v / (1.0f + expf(-v));
      }
      xreg[u] = v;
    }
  };
  fetch(0);
  for (int kc = 0; kc < K; kc += KC) {
#pragma unroll
    for (int u = 0; u < OPT; ++u) {
      const int idx = threadIdx.x + u * 256, r = idx >> 3, c4 = idx & 7;
      *reinterpret_cast<float4*>(&s_w[r * LDW + c4 * 4]) = wreg[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = threadIdx.x + u * 256, r = idx >> 5, c = idx & 31;
      s_in[c * 32 + r] = xreg[u];
    }
    __syncthreads();
    if (kc + KC < K) fetch(kc + KC);
#pragma unroll
    for (int k = 0; k < KC; k += 4) {
      float4 x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = *reinterpret_cast<const float4*>(&s_in[(k + q) * 32 + ty * 4]);
#pragma unroll
      for (int i = 0; i < OPT; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(&s_w[(tx + 32 * i) * LDW + k]);
        const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[0][i] = fmaf(x[q].x, wk[q], acc[0][i]);
          acc[1][i] = fmaf(x[q].y, wk[q], acc[1][i]);
          acc[2][i] = fmaf(x[q].z, wk[q], acc[2][i]);
          acc[3][i] = fmaf(x[q].w, wk[q], acc[3][i]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = ty * 4 + j;
    if (r >= rows) continue;
    const int n = n0 + r;
    int64_t cls = -1;
    if (label_emb != nullptr && classes != nullptr) cls = classes[n % Ncls];
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
      const int o = o0 + tx + 32 * i;
      if (o < O) {
        float v = acc[j][i] + bias[o];
        if (cls >= 0) v += label_emb[static_cast<size_t>(cls) * O + o];   // null class (-1) contributes zero (adm.py:551-553)
        out[static_cast<size_t>(n) * O + o] = v;
      }
    }
  }
}

// Small-O variant (time_embed, O ~ 1K): one warp per output feature, lanes across K, all (<= 32) batch rows kept as
// per-lane partial sums and reduced by shuffles.  The whole weight row of a feature is in flight at once (K <= 1024),
// so the two dependent time_embed Linears are not a chain of latency-bound K chunks.  K % 128 == 0.
template <int KV>   // KV = K / 128
__global__ void __launch_bounds__(256) linear_warp_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out, int N,
                                                          int O, int silu_in, const float* __restrict__ label_emb,
                                                          const int64_t* __restrict__ classes, int Ncls) {
  constexpr int K = KV * 128;
  extern __shared__ __align__(16) float s_x[];      // [32 rows][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * 32;
  const int rows = min(32, N - n0);
  const int o = blockIdx.x * 8 + warp;
  float4 w[KV];
  if (o < O) {
#pragma unroll
    for (int v = 0; v < KV; ++v) w[v] = ldg_f4(W + static_cast<size_t>(o) * K + v * 128 + lane * 4);
  }
  for (int idx = threadIdx.x; idx < 32 * K; idx += 256) {
    const int r = idx / K;
    float x = 0.f;
    if (r < rows) {
      x = __ldg(in + static_cast<size_t>(n0) * K + idx);
      if (silu_in) x = x / (1.0f + expf(-x));
    }
    s_x[idx] = x;
  }
  __syncthreads();
  if (o >= O) return;
  float mine = 0.f;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
      const float4 x = *reinterpret_cast<const float4*>(&s_x[r * K + v * 128 + lane * 4]);
      a = fmaf(x.x, w[v].x, a); a = fmaf(x.y, w[v].y, a); a = fmaf(x.z, w[v].z, a); a = fmaf(x.w, w[v].w, a);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == r) mine = a;
  }
  if (lane < rows) {
    const int n = n0 + lane;
    float v = mine + bias[o];
    if (label_emb != nullptr && classes != nullptr) {
      const int64_t cls = classes[n % Ncls];
      if (cls >= 0) v += label_emb[static_cast<size_t>(cls) * O + o];
    }
    out[static_cast<size_t>(n) * O + o] = v;
  }
}

// x_t[nb][k][r] = act(in[nb*32 + r][k]) (zero rows past N): the k-major operand layout of film_table_kernel
__global__ void silu_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int K, int silu_in) {
  const int total = ((N + 31) / 32) * K * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx & 31, k = (idx >> 5) % K, nb = idx / (32 * K);
    const int n = nb * 32 + r;
    float v = 0.f;
    if (n < N) {
      v = __ldg(in + static_cast<size_t>(n) * K + k);
      if (silu_in) v = v / (1.0f + expf(-v));
    }
    out[idx] = v;
  }
}

// FiLM table: out[N][O] = x[N][K] * W^T + bias with O ~ 40K, K ~ 1K: ~170 MB of fp32 weights per forward, read once.
// Weight-bandwidth bound, so the weights are streamed by 1-D bulk async copies through an 8-stage shared-memory ring
// (160 KB in flight per SM) instead of register prefetch.  Packed layout (made by the weight packer):
//   Wp[K/32][O][32], the eight 16-byte groups of each 32-float row XOR-swizzled with (o & 7), so that a 128-output x 32-k
//   tile is one contiguous 16 KB run in HBM and the consumers' 128-bit shared-memory reads are conflict-free.
// Block = 8 compute warps (thread = 4 batch rows x 4 output features) + 1 producer warp; blocks stride over 128-output
// tiles, the ring keeps running across tiles.
struct FilmCfg {
  static constexpr int TO = 128, KC = 32, STAGES = 8;
  static constexpr int W_BYTES = TO * KC * 4, X_BYTES = KC * 32 * 4, STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int THREADS = 288;
};

__global__ void __launch_bounds__(FilmCfg::THREADS, 1) film_table_kernel(const float* __restrict__ Wp, const float* __restrict__ x_t,
                                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                                         int N, int K, int O) {
  using Cfg = FilmCfg;
  extern __shared__ __align__(128) uint8_t film_smem[];
  uint8_t* ring = film_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(film_smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = blockIdx.y;
  const int tiles = (O + Cfg::TO - 1) / Cfg::TO;
  const int kchunks = K / Cfg::KC;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) {
      uint32_t seq = 0;
      const float* xb = x_t + static_cast<size_t>(nb) * K * 32;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int o0 = t * Cfg::TO;
        const uint32_t wbytes = static_cast<uint32_t>(min(Cfg::TO, O - o0)) * Cfg::KC * 4;
        for (int kc = 0; kc < kchunks; ++kc, ++seq) {
          const uint32_t s = seq % Cfg::STAGES, ph = (seq / Cfg::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = ring + s * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], wbytes + Cfg::X_BYTES);
          bulk_load_1d(st, Wp + (static_cast<size_t>(kc) * O + o0) * Cfg::KC, wbytes, &full[s]);
          bulk_load_1d(st + Cfg::W_BYTES, xb + static_cast<size_t>(kc) * Cfg::KC * 32, Cfg::X_BYTES, &full[s]);
        }
      }
    }
    return;
  }
  const int tx = lane, ty = warp;
  uint32_t seq = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int o0 = t * Cfg::TO;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    for (int kc = 0; kc < kchunks; ++kc, ++seq) {
      const uint32_t s = seq % Cfg::STAGES, ph = (seq / Cfg::STAGES) & 1;
      mbar_wait(&full[s], ph);
      const float4* sw = reinterpret_cast<const float4*>(ring + s * Cfg::STAGE_BYTES);
      const float4* sx = reinterpret_cast<const float4*>(ring + s * Cfg::STAGE_BYTES + Cfg::W_BYTES);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = sx[(k4 * 4 + q) * 8 + ty];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w = sw[(tx + 32 * i) * 8 + (k4 ^ (tx & 7))];
          const float wk[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[0][i] = fmaf(x[q].x, wk[q], acc[0][i]);
            acc[1][i] = fmaf(x[q].y, wk[q], acc[1][i]);
            acc[2][i] = fmaf(x[q].z, wk[q], acc[2][i]);
            acc[3][i] = fmaf(x[q].w, wk[q], acc[3][i]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nb * 32 + ty * 4 + j;
      if (n >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int o = o0 + tx + 32 * i;
        if (o < O) out[static_cast<size_t>(n) * O + o] = acc[j][i] + bias[o];
      }
    }
  }
}

}  // namespace ivid
