// Micro-benchmark of the fp16 -> fp16 GroupNorm-apply stream (the largest gn_apply shape of the L model) with a few
// kernel variants, to find out what limits it below the copy bandwidth.  Build: see tools/micro/build.sh
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../ivid_b200/csrc/common.cuh"
#include "../../ivid_b200/csrc/elementwise.cuh"
using namespace ivid;

struct P { const __half* x; __half* y; const float* ab; int C, HW, ppb; };

__device__ __forceinline__ float silu_tanh(float x) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x)); return fmaf(0.5f * x, t, 0.5f * x); }

template <int U, bool HOIST, int MINB, bool STREAM, bool TANH = false>
__global__ void __launch_bounds__(256, MINB) k_apply(const P p) {
  extern __shared__ float s_ab[];
  const int C = p.C, c8 = C >> 3, n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += 256) {
    s_ab[(c & 7) * c8 + (c >> 3)] = p.ab[(n * C + c) * 2];
    s_ab[C + (c & 7) * c8 + (c >> 3)] = p.ab[(n * C + c) * 2 + 1];
  }
  __syncthreads();
  const int pix0 = blockIdx.x * p.ppb;
  const int npix = min(p.ppb, p.HW - pix0);
  const int items = npix * c8;
  float A[8], B[8];
  if (HOIST) {
    const int cg = threadIdx.x % c8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { A[j] = s_ab[j * c8 + cg]; B[j] = s_ab[C + j * c8 + cg]; }
  }
  const __half* xb = p.x + (static_cast<size_t>(n) * p.HW + pix0) * C;
  __half* yb = p.y + (static_cast<size_t>(n) * p.HW + pix0) * C;
  for (int it0 = threadIdx.x; it0 + (U - 1) * 256 < items; it0 += U * 256) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint4* src = reinterpret_cast<const uint4*>(xb + static_cast<size_t>(it0 + u * 256) * 8);
      if (STREAM) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(raw[u].x), "=r"(raw[u].y), "=r"(raw[u].z), "=r"(raw[u].w) : "l"(src));
      else raw[u] = __ldg(src);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * 256;
      const int cg = it % c8;
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u]);
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        float y0, y1;
        if (HOIST) { y0 = fmaf(f.x, A[2 * j], B[2 * j]); y1 = fmaf(f.y, A[2 * j + 1], B[2 * j + 1]); }
        else { y0 = fmaf(f.x, s_ab[(2 * j) * c8 + cg], s_ab[C + (2 * j) * c8 + cg]); y1 = fmaf(f.y, s_ab[(2 * j + 1) * c8 + cg], s_ab[C + (2 * j + 1) * c8 + cg]); }
        if (TANH) { y0 = silu_tanh(y0); y1 = silu_tanh(y1); } else { y0 = silu_f(y0); y1 = silu_f(y1); }
        pk[j] = pack_h2(y0, y1);
      }
      uint4* dst = reinterpret_cast<uint4*>(yb + static_cast<size_t>(it) * 8);
      if (STREAM) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      else *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

__global__ void k_copy(const uint4* __restrict__ x, uint4* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) y[i] = __ldg(x + i);
}

__global__ void k_fill(uint4* __restrict__ y, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) y[i] = v;
}
__global__ void k_read(const uint4* __restrict__ x, uint4* __restrict__ y, size_t n) {
  uint4 a = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(x + i); a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
  }
  if (a.x == 0x12345678u) y[0] = a;
}

template <typename F>
static float time_it(F f, int reps = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const int N = 32, HW = 128 * 128, C = argc > 1 ? atoi(argv[1]) : 256;
  const size_t el = static_cast<size_t>(N) * HW * C;
  __half *x, *y; float* ab;
  cudaMalloc(&x, el * 2); cudaMalloc(&y, el * 2); cudaMalloc(&ab, N * C * 8);
  cudaMemset(x, 0x3c, el * 2); cudaMemset(ab, 0, N * C * 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const double gb = el * 4.0 / 1e9;
  auto run = [&](const char* name, auto kern, int blocks_per_sm, int waves) {
    P p{x, y, ab, C, HW, 0};
    const int bpn = std::max(1, sms * blocks_per_sm * waves / N);
    p.ppb = (HW + bpn - 1) / bpn;
    dim3 grid((HW + p.ppb - 1) / p.ppb, N);
    const float ms = time_it([&] { kern<<<grid, 256, C * 8>>>(p); });
    printf("%-34s grid %4dx%d ppb %5d : %.4f ms  %.0f GB/s  (%s)\n", name, grid.x, grid.y, p.ppb, ms, gb / ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  {
    const float ms = time_it([&] { k_copy<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), el / 8); });
    printf("%-34s : %.4f ms  %.0f GB/s\n", "plain copy (16B/thread grid-stride)", ms, gb / ms * 1e3);
    const float msf = time_it([&] { k_fill<<<sms * 8, 256>>>(reinterpret_cast<uint4*>(y), el / 8); });
    printf("%-34s : %.4f ms  %.0f GB/s\n", "pure write (16B/thread)", msf, el * 2.0 / 1e9 / msf * 1e3);
    const float msm = time_it([&] { cudaMemsetAsync(y, 0, el * 2); });
    printf("%-34s : %.4f ms  %.0f GB/s\n", "cudaMemset", msm, el * 2.0 / 1e9 / msm * 1e3);
    const float msr = time_it([&] { k_read<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), el / 8); });
    printf("%-34s : %.4f ms  %.0f GB/s\n", "pure read (16B/thread)", msr, el * 2.0 / 1e9 / msr * 1e3);
    const float ms2 = time_it([&] { cudaMemcpyAsync(y, x, el * 2, cudaMemcpyDeviceToDevice); });
    printf("%-34s : %.4f ms  %.0f GB/s\n", "cudaMemcpy D2D", ms2, gb / ms2 * 1e3);
  }
  {
    // the production kernel on the same stream (statistics prologue, generic indexing)
    double* st; float *gamma, *beta, *film; __half* x1;
    cudaMalloc(&st, N * C * 16); cudaMalloc(&gamma, C * 8); cudaMalloc(&beta, C * 8); cudaMalloc(&film, N * C * 2 * 8); cudaMalloc(&x1, el * 2);
    std::vector<double> hst(N * C * 2);
    for (int i = 0; i < N * C; ++i) { hst[2 * i] = 0.1 * HW; hst[2 * i + 1] = 1.5 * HW; }
    cudaMemcpy(st, hst.data(), hst.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(gamma, 0, C * 8); cudaMemset(beta, 0, C * 8); cudaMemset(film, 0, N * C * 2 * 8); cudaMemset(x1, 0x3c, el * 2);
    for (int concat = 0; concat < 2; ++concat)
      for (int waves = 1; waves <= 4; waves *= 2) {
        GnApplyParams q{};
        q.x0h = x; q.x1h = concat ? x1 : nullptr; q.C0 = C; q.C1 = concat ? C : 0; q.N = N; q.H = 128; q.W = 128; q.mode = 0; q.silu = 1;
        q.stats0 = st; q.stats1 = st; q.groups = 32; q.inv_count = 1.0 / HW; q.eps = 1e-5f; q.gamma = gamma; q.beta = beta;
        q.film = concat ? nullptr : film; q.film_ld = 2 * C; q.film_off = 0;
        __half* yy; cudaMalloc(&yy, el * 2 * (1 + concat));
        q.out_act = yy;
        const int bpn = std::max(1, sms * 3 * waves / N);
        q.pix_per_block = (HW + bpn - 1) / bpn;
        dim3 grid((HW + q.pix_per_block - 1) / q.pix_per_block, N);
        const int Ct = q.C0 + q.C1;
        const float ms = time_it([&] {
          if (256 % (Ct / 8) == 0) gn_apply_h16_kernel<true><<<grid, 256, Ct * 8>>>(q); else gn_apply_h16_kernel<false><<<grid, 256, Ct * 8>>>(q);
        });
        printf("production h16 kernel concat=%d waves=%d grid %dx%d : %.4f ms  %.0f GB/s (%s)\n", concat, waves, grid.x, grid.y, ms, gb * (1 + concat) / ms * 1e3,
               cudaGetErrorString(cudaGetLastError()));
        cudaFree(yy);
      }
  }
  run("U8 lds 3/SM 1 wave (current)", k_apply<8, false, 3, false>, 3, 1);
  run("U8 hoist 3/SM 1 wave", k_apply<8, true, 3, false>, 3, 1);
  run("U8 hoist 3/SM 4 waves", k_apply<8, true, 3, false>, 3, 4);
  run("U8 hoist tanh 3/SM 1 wave", k_apply<8, true, 3, false, true>, 3, 1);
  run("U8 hoist tanh 3/SM 4 waves", k_apply<8, true, 3, false, true>, 3, 4);
  run("U8 hoist stream 3/SM", k_apply<8, true, 3, true>, 3, 1);
  run("U4 hoist 4/SM", k_apply<4, true, 4, false>, 4, 1);
  run("U4 hoist 6/SM", k_apply<4, true, 6, false>, 6, 1);
  run("U4 hoist stream 6/SM", k_apply<4, true, 6, true>, 6, 1);
  run("U16 hoist 2/SM", k_apply<16, true, 2, false>, 2, 1);
  run("U2 hoist 8/SM", k_apply<2, true, 8, false>, 8, 1);
  run("U8 lds 3/SM 4 waves", k_apply<8, false, 3, false>, 3, 4);
  return 0;
}
