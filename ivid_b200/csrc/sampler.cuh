// Fused denoising-step kernels: classifier-free-guidance mix + DDPM / DDIM update (+ multiview replace/constrain
// guidance) in ONE HBM pass over [N,4,H,W], coefficients read from a device-resident table (no per-step H2D).
//   reference: ClassifierFreeGuidance.model_inference classifier_free_guidance.py:39-42
//              DdpmSampler.p_mean_variance / sample_once   samplers/ddpm.py:85-100,127-131
//              DdimSampler.sample_once                      samplers/ddim.py:81-103
//              InpaintCFG.make_cond_inputs                  frameworks/inpaint_cfg.py:33-49
//              SuperResCFG.make_cond_inputs                 frameworks/sr_cfg.py:31-36
#pragma once
#include "common.cuh"

namespace ivid {

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG + Box-Muller (in-kernel N(0,1) for the production path; parity tests inject noise).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x) + 0.5f) * 2.3283064365386963e-10f; }
// four N(0,1) draws for (seed, stream, idx)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t stream, uint32_t idx) {
  const uint4 r = philox4x32_10(make_uint4(idx, stream, 0u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float r0 = sqrtf(-2.0f * logf(u01(r.x))), a0 = 6.283185307179586f * u01(r.y);
  const float r1 = sqrtf(-2.0f * logf(u01(r.z))), a1 = 6.283185307179586f * u01(r.w);
  float s0, c0, s1, c1;
  sincosf(a0, &s0, &c0);
  sincosf(a1, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// Per-timestep coefficient row (fp32 casts of the reference's float64 numpy tables).
struct StepCoef {
  float sqrt_recip_acp;      // sqrt(1/acp[t])
  float sqrt_recipm1_acp;    // sqrt(1/acp[t] - 1)
  float post_mean_coef1;     // DDPM
  float post_mean_coef2;     // DDPM
  float post_logvar;         // DDPM posterior_log_variance_clipped[t]
  float acp;                 // alphas_cumprod[t]
  float acp_prev;            // alphas_cumprod_prev[t]  (acp_prev[0] = 1)
  float pad;
};

struct GuideParams {           // DDIM multiview guidance (all maps fp32 NCHW, nullptr = disabled)
  const float* rgb;            // [N,3,H,W]
  const float* rgb_mask;       // [N,1,H,W]
  const float* depth;          // [N,1,H,W]
  const float* depth_mask;     // [N,1,H,W]
  const float* convex;         // [N,1,H,W]
  float w_rgb, w_rgb_c;        // w and (1-w) as the reference's python floats cast to fp32
  float w_depth, w_depth_c;
  float w_convex, w_convex_c;
};

struct StepParams {
  const float* x_t;            // [N,C,H,W]
  const float* eps;            // [2N or N,C,H,W]: rows [0,N) conditional, [N,2N) unconditional when cfg
  const float* noise;          // [N,C,H,W] injected N(0,1) or nullptr -> Philox
  float* x_prev;               // [N,C,H,W]
  float* pred_x0;              // optional
  const StepCoef* table;       // [T]
  const int* t_index;          // device scalar: table row for the model timestep (t for DDPM, t-1 for DDIM)
  const int* t_prev;           // DDIM: device scalar t_prev (acp_prev row); DDPM: unused
  int N, C, HW;
  int cfg;                     // 1: eps = (1+s)*eps_c - s*eps_u;  2: eps = (1+s)*eps_c (strength <= 0: no null-class forward)
  float strength;
  int clip;
  float eta;
  uint64_t seed;
  uint32_t stream;             // Philox stream id (step counter supplied by caller) when noise == nullptr
  const int* stream_dev;       // optional device step counter added to `stream`
  GuideParams g;
};

// (1 + strength) * eps_c - strength * eps_u (only evaluated with strength > 0).  Every product / sum of the step arithmetic is
// written with explicit round-to-nearest intrinsics (no compiler-chosen FMA contraction), so that step_kernel and
// head_step_kernel - the same formulas inlined into different kernels - produce identical bits.
__device__ __forceinline__ float cfg_mix(float ec, float eu, float strength) {
  return __fsub_rn(__fmul_rn(__fadd_rn(1.0f, strength), ec), __fmul_rn(strength, eu));
}
__device__ __forceinline__ float mix_eps(const StepParams& p, size_t i, size_t total) {
  const float ec = p.eps[i];
  if (!p.cfg) return ec;
  if (p.cfg == 2) return __fmul_rn(__fadd_rn(1.0f, p.strength), ec);
  return cfg_mix(ec, p.eps[total + i], p.strength);
}

// classifier-free-guidance mix alone (framework.model_inference): out = (1+s)*eps[0:n) - s*eps[n:2n)
__global__ void __launch_bounds__(256) cfg_mix_kernel(const float* __restrict__ eps2, float* __restrict__ out, size_t n4, float strength) {
  const float4* a = reinterpret_cast<const float4*>(eps2);
  const float4* b = a + n4;
  float4* o = reinterpret_cast<float4*>(out);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 c = __ldg(a + i), u = __ldg(b + i);
    o[i] = make_float4((1.0f + strength) * c.x - strength * u.x, (1.0f + strength) * c.y - strength * u.y,
                       (1.0f + strength) * c.z - strength * u.z, (1.0f + strength) * c.w - strength * u.w);
  }
}

// Per-step scalars derived once per thread from the device-resident step state.
struct StepScalars {
  StepCoef k;
  float nz;          // DDPM: t != 0 ; DDIM: t_prev != 0
  float sd;          // DDPM: exp(0.5 * posterior_log_variance_clipped)
  float sigma, c_x0, c_eps;    // DDIM
  uint32_t stream;
};
template <bool kDdim>
__device__ __forceinline__ StepScalars step_scalars(const StepParams& p) {
  StepScalars s;
  s.k = p.table[*p.t_index];
  s.stream = p.stream + (p.stream_dev ? static_cast<uint32_t>(*p.stream_dev) : 0u);
  if (!kDdim) {
    s.nz = (*p.t_index != 0) ? 1.0f : 0.0f;
    s.sd = expf(__fmul_rn(0.5f, s.k.post_logvar));
    s.sigma = 0.f; s.c_x0 = 0.f; s.c_eps = 0.f;
  } else {
    const int tprev = *p.t_prev;
    s.nz = (tprev != 0) ? 1.0f : 0.0f;
    const float ab = s.k.acp;
    const float abp = (tprev == 0) ? 1.0f : p.table[tprev - 1].acp;   // alphas_cumprod_prev[t_prev]
    s.sigma = __fmul_rn(__fmul_rn(p.eta, sqrtf(__fdiv_rn(__fsub_rn(1.0f, abp), __fsub_rn(1.0f, ab)))), sqrtf(__fsub_rn(1.0f, __fdiv_rn(ab, abp))));
    s.c_x0 = sqrtf(abp);
    s.c_eps = sqrtf(__fsub_rn(__fsub_rn(1.0f, abp), __fmul_rn(s.sigma, s.sigma)));
    s.sd = 0.f;
  }
  return s;
}
// x_{t-1} and x_0 of one element (sample n, channel c, pixel pix) from x_t, the (already guidance-mixed) eps and the N(0,1) draw z
template <bool kDdim>
__device__ __forceinline__ void step_element(const StepParams& p, const StepScalars& s, int n, int c, size_t pix, float xt, float e, float z,
                                             float& xo, float& x0o) {
  const StepCoef& k = s.k;
  auto mul = [](float a, float b) { return __fmul_rn(a, b); };
  auto add = [](float a, float b) { return __fadd_rn(a, b); };
  auto sub = [](float a, float b) { return __fsub_rn(a, b); };
  float x0 = sub(mul(k.sqrt_recip_acp, xt), mul(k.sqrt_recipm1_acp, e));
  if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  if (!kDdim) {
    const float mean = add(mul(k.post_mean_coef1, x0), mul(k.post_mean_coef2, xt));
    xo = add(mean, mul(mul(s.nz, s.sd), z));
    x0o = x0;
    return;
  }
  if (c < 3) {
    if (p.g.rgb != nullptr) {
      const float y = p.g.rgb[(static_cast<size_t>(n) * 3 + c) * p.HW + pix];
      const float m = p.g.rgb_mask[static_cast<size_t>(n) * p.HW + pix];
      x0 = add(mul(sub(1.0f, s.nz), x0), mul(s.nz, add(mul(add(mul(p.g.w_rgb, y), mul(p.g.w_rgb_c, x0)), m), mul(x0, sub(1.0f, m)))));
    }
  } else if (p.g.depth != nullptr) {
    const float y = p.g.depth[static_cast<size_t>(n) * p.HW + pix];
    const float m = p.g.depth_mask[static_cast<size_t>(n) * p.HW + pix];
    x0 = add(mul(add(mul(p.g.w_depth, y), mul(p.g.w_depth_c, x0)), m), mul(x0, sub(1.0f, m)));
    if (p.g.convex != nullptr) {
      const float cv = p.g.convex[static_cast<size_t>(n) * p.HW + pix];
      x0 = add(mul(x0, m), mul(add(mul(p.g.w_convex, fmaxf(x0, cv)), mul(p.g.w_convex_c, x0)), sub(1.0f, m)));
    }
  }
  const float e2 = __fdiv_rn(sub(mul(k.sqrt_recip_acp, xt), x0), k.sqrt_recipm1_acp);
  const float mean = add(mul(s.c_x0, x0), mul(s.c_eps, e2));
  xo = add(mean, mul(mul(s.nz, s.sigma), z));
  x0o = x0;
}
// the four N(0,1) draws of elements [i, i+4) of the flattened [N,C,H,W] tensor (i % 4 == 0): injected or Philox(seed, stream, i/4)
__device__ __forceinline__ void step_noise4(const StepParams& p, const StepScalars& s, size_t i, bool needed, float (&z)[4]) {
  z[0] = z[1] = z[2] = z[3] = 0.f;
  if (!needed) return;
  const float4 t = p.noise != nullptr ? ldg_f4(p.noise + i) : philox_normal4(p.seed, s.stream, static_cast<uint32_t>(i >> 2));
  z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
}

// one thread = 4 consecutive pixels of one (n, c) plane
template <bool kDdim>
__global__ void __launch_bounds__(256) step_kernel(const StepParams p) {
  const size_t total = static_cast<size_t>(p.N) * p.C * p.HW;
  const StepScalars s = step_scalars<kDdim>(p);
  for (size_t i4 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i4 * 4 < total;
       i4 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t i = i4 * 4;
    const int plane = static_cast<int>(i / p.HW);      // n*C + c
    const int n = plane / p.C, c = plane % p.C;
    const size_t pix = i - static_cast<size_t>(plane) * p.HW;
    float z[4];
    step_noise4(p, s, i, !kDdim || s.sigma != 0.0f, z);
    float xo[4], x0o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) step_element<kDdim>(p, s, n, c, pix + j, p.x_t[i + j], mix_eps(p, i + j, total), z[j], xo[j], x0o[j]);
    stg_f4(p.x_prev + i, make_float4(xo[0], xo[1], xo[2], xo[3]));
    if (p.pred_x0) stg_f4(p.pred_x0 + i, make_float4(x0o[0], x0o[1], x0o[2], x0o[3]));
  }
}

// Output head + denoising step in ONE kernel (the last node of the forward's CUDA graph): eps of both guidance halves is formed
// from the tap columns Y of the output head's 1x1 GEMM (the shift-and-add of eps_gather_kernel, same summation order), mixed, and
// pushed through the DDPM / DDIM update without ever being written to HBM.  One thread = 4 consecutive pixels (one row segment)
// of one sample, all Co = 4 channels; noise indices and arithmetic are those of step_kernel, so both routes agree bit for bit.
struct HeadStepParams {
  StepParams sp;
  const float* Y;          // [2N or N][H][W][ldy] tap columns (tap*Co + c)
  const float* bias;       // [Co]
  int H, W, ldy;
};
__device__ __forceinline__ void head_eps4(const HeadStepParams& h, int n, int y, int x, float (&e)[4]) {
  e[0] = e[1] = e[2] = e[3] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int hh = y + tap / 3 - 1, ww = x + tap % 3 - 1;
    if (hh < 0 || hh >= h.H || ww < 0 || ww >= h.W) continue;
    const float4 v = ldg_f4(h.Y + ((static_cast<size_t>(n) * h.H + hh) * h.W + ww) * h.ldy + tap * 4);
    e[0] += v.x; e[1] += v.y; e[2] += v.z; e[3] += v.w;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) e[c] += __ldg(h.bias + c);
}
template <bool kDdim>
__global__ void __launch_bounds__(256) head_step_kernel(const HeadStepParams h) {
  const StepParams& p = h.sp;
  const StepScalars s = step_scalars<kDdim>(p);
  const int w4 = h.W / 4;
  const size_t groups = static_cast<size_t>(p.N) * h.H * w4;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < groups; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(g % w4);
    const int y = static_cast<int>((g / w4) % h.H);
    const int n = static_cast<int>(g / (static_cast<size_t>(w4) * h.H));
    float e[4][4];                  // [pixel][channel]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ec[4];
      head_eps4(h, n, y, xg * 4 + j, ec);
      if (p.cfg == 1) {
        float eu[4];
        head_eps4(h, n + p.N, y, xg * 4 + j, eu);
#pragma unroll
        for (int c = 0; c < 4; ++c) ec[c] = cfg_mix(ec[c], eu[c], p.strength);
      } else if (p.cfg == 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) ec[c] = __fmul_rn(__fadd_rn(1.0f, p.strength), ec[c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) e[j][c] = ec[c];
    }
    const size_t pix = static_cast<size_t>(y) * h.W + xg * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t i = (static_cast<size_t>(n) * 4 + c) * p.HW + pix;
      float z[4];
      step_noise4(p, s, i, !kDdim || s.sigma != 0.0f, z);
      const float4 xt = ldg_f4(p.x_t + i);
      const float xtv[4] = {xt.x, xt.y, xt.z, xt.w};
      float xo[4], x0o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) step_element<kDdim>(p, s, n, c, pix + j, xtv[j], e[j][c], z[j], xo[j], x0o[j]);
      stg_f4(p.x_prev + i, make_float4(xo[0], xo[1], xo[2], xo[3]));
      if (p.pred_x0) stg_f4(p.pred_x0 + i, make_float4(x0o[0], x0o[1], x0o[2], x0o[3]));
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Conditional-model input assembly, fused with the NCHW fp32 -> NHWC fp16(64 ch) packing.
// ----------------------------------------------------------------------------------------------
struct CondPackParams {
  const float* x;         // [Nx,4,H,W]
  const float* y;         // inpaint: [Nx,4,H,W] warped RGBD;  superres: [Nx,4,H/2,W/2] low-res RGBD
  const float* mask;      // inpaint [Nx,1,H,W]
  const float* mask_rgb;  // inpaint [Nx,1,H,W] or nullptr (then mask is used and no mask_rgb channel is emitted)
  const float* noise;     // inpaint: injected [Nx,4,H,W] (rgb noise 3 + depth noise 1) or nullptr -> Philox
  __half* out;            // [N,H,W,64]
  int N, Nx, H, W;
  int kind;               // 1 = InpaintCFG, 2 = SuperResCFG
  uint64_t seed;
  uint32_t stream;
  const int* stream_dev;
};

__global__ void __launch_bounds__(256) cond_pack_kernel(const CondPackParams p) {
  // phase 1: one thread per pixel assembles the Cin (<= 16) fp32 input channels into shared memory;
  // phase 2: the block writes the 64 fp16 operand channels (two-term split hi | lo | hi, see pack_input_kernel) with
  //          16-byte coalesced stores, one thread per (pixel, 8-channel group).
  __shared__ float s_ch[256][17];
  const int HW = p.H * p.W;
  const size_t total = static_cast<size_t>(p.N) * HW;
  const uint32_t stream = p.stream + (p.stream_dev ? static_cast<uint32_t>(*p.stream_dev) : 0u);
  const int Cin = p.kind == 1 ? (p.mask_rgb ? 10 : 9) : 8;
  for (size_t base = blockIdx.x * static_cast<size_t>(blockDim.x); base < total; base += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t idx = base + threadIdx.x;
    if (idx < total) {
      const int n = static_cast<int>(idx / HW) % p.Nx;
      const int pix = static_cast<int>(idx % HW);
      float ch[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) ch[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) ch[c] = p.x[(static_cast<size_t>(n) * 4 + c) * HW + pix];
      if (p.kind == 1) {
        const float m = p.mask[static_cast<size_t>(n) * HW + pix];
        const float mr = p.mask_rgb ? p.mask_rgb[static_cast<size_t>(n) * HW + pix] : m;
        float z[4];
        if (p.noise != nullptr) {
#pragma unroll
          for (int c = 0; c < 4; ++c) z[c] = p.noise[(static_cast<size_t>(n) * 4 + c) * HW + pix];
        } else {
          const float4 t = philox_normal4(p.seed, stream, static_cast<uint32_t>(static_cast<size_t>(n) * HW + pix));
          z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
        }
        int o = 4;
        if (p.mask_rgb) ch[o++] = mr;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ch[o++] = p.y[(static_cast<size_t>(n) * 4 + c) * HW + pix] * mr + z[c] * (1.0f - mr);
        ch[o++] = p.y[(static_cast<size_t>(n) * 4 + 3) * HW + pix] * m + z[3] * (1.0f - m);
        ch[o++] = m;
      } else {
        // bilinear 2x upsample, align_corners=False: src = (dst + 0.5)/2 - 0.5, clamped at 0 (ATen upsample_bilinear2d)
        const int h = pix / p.W, w = pix % p.W;
        const int Hs = p.H / 2, Ws = p.W / 2;
        float sy = (h + 0.5f) * 0.5f - 0.5f; if (sy < 0.f) sy = 0.f;
        float sx = (w + 0.5f) * 0.5f - 0.5f; if (sx < 0.f) sx = 0.f;
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* s = p.y + (static_cast<size_t>(n) * 4 + c) * Hs * Ws;
          const float v = (1.f - ly) * ((1.f - lx) * s[y0 * Ws + x0] + lx * s[y0 * Ws + x1]) +
                          ly * ((1.f - lx) * s[y1 * Ws + x0] + lx * s[y1 * Ws + x1]);
          ch[4 + c] = v;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s_ch[threadIdx.x][j] = ch[j];
    }
    __syncthreads();
    const int live = static_cast<int>(min(static_cast<size_t>(256), total - base));
    for (int it = threadIdx.x; it < live * 8; it += 256) {
      const int px = it >> 3, g = it & 7;
      uint32_t w4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __half e[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int oc = g * 8 + 2 * k + q;
          const int seg = oc / Cin;
          e[q] = seg < 3 ? split_term(s_ch[px][oc - seg * Cin], seg) : __float2half_rn(0.f);
        }
        const __half2 h2 = __halves2half2(e[0], e[1]);
        w4[k] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(p.out + (base + px) * 64 + g * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
    __syncthreads();
  }
}

}  // namespace ivid
