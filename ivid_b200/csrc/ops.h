// Host-side launch descriptors for the sm_100a kernels (built once at plan time, replayed every step).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_util.h"

namespace ivid {

// mirrors of the device parameter structs (defined in the .cuh files; redeclared opaque here through includes in ops.cu)
struct ConvLaunch;
struct AttnLaunch;

struct ConvDesc {
  const void* act0 = nullptr; int C0 = 0; int taps0 = 9;   // segment 0: fp16 NHWC activation, 9 = 3x3, 1 = 1x1
  const void* act1 = nullptr; int C1 = 0; int taps1 = 1;   // optional segment 1 (1x1 skip over another tensor)
  const void* act2 = nullptr; int C2 = 0; int taps2 = 1;   // optional segment 2 (second half of a virtual concat)
  void* out16 = nullptr;                                   // optional fp16 NHWC copy of an fp32 output (same ldc)
  const void* weight = nullptr;                            // fp16 [cout_pad][Ktot], Ktot = taps0*C0 + taps1*C1
  int cout_pad = 0;
  int cout = 0;                                            // valid output channels
  const float* bias = nullptr;                             // [cout_pad]
  const float* residual = nullptr; int ldr = 0;            // fp32 NHWC
  bool residual_up = false;                                // residual is [N][H/2][W/2][ldr]: added through a nearest-2x upsample (conv_can_res_up)
  void* out = nullptr; int ldc = 0; int out_mode = 0;      // 0 fp32 NHWC, 1 fp16 NHWC, 2 fp32 NCHW
  double* stats = nullptr;                                 // optional fused GroupNorm statistics of the output [N][cout][2]
  int N = 0, H = 0, W = 0;
  // fold mode (conv_fold_ok): the 3x3 segments are RAW fp16 tensors and y = silu(A x + B) is applied in the kernel's operand path
  const void* fold_ab = nullptr;                           // float2 [N][fold_C] from launch_gn_coeff
  int fold_C = 0;
  int fold_off0 = -1, fold_off1 = -1, fold_off2 = -1;      // channel offset of each segment in the table (< 0: not transformed)
};
// whether a 3x3 conv of this shape runs on the kernel variant that can fold a GroupNorm apply into its operand path
// (CTA pairs, 8 x 16-pixel tiles; IVID_FOLD=1 opts in)
bool conv_fold_ok(int N, int H, int W, int cout_pad, bool residual_up);

// opaque, heap-allocated launch records (hold the CUtensorMaps)
ConvLaunch* conv_launch_create(const ConvDesc& d);
void conv_launch_destroy(ConvLaunch* l);
void conv_launch_run(const ConvLaunch* l, cudaStream_t s);
int conv_launch_bn(const ConvLaunch* l);
void conv_launch_run_out(const ConvLaunch* l, void* out, cudaStream_t s);   // same launch, output pointer overridden
int conv_pick_bn(int cout_pad, int m_tiles = 0);          // m_tiles > 0: wave-quantisation aware choice
// whether a conv with this many (unpadded) output channels can also emit the fp16 copy of its fp32 NHWC output
bool conv_can_out16(int cout);
// whether the conv epilogue can add a half-resolution residual through a nearest-2x upsample (output width W, Cout)
bool conv_can_res_up(int W, int cout);
bool conv_can_fuse_stats(int H, int W);                    // epilogue statistics need >= 32 pixels of one sample per warp
int conv_pad_cout(int cout);

AttnLaunch* attn_launch_create(const void* qkv, int N, int T, int C, void* out);
void attn_launch_destroy(AttnLaunch* l);
void attn_launch_run(const AttnLaunch* l, cudaStream_t s);

void launch_gn_stats(const float* x, double* stats, int N, int HW, int C, cudaStream_t s);
struct GnApplyDesc {
  const void* x0 = nullptr; const void* x1 = nullptr; int C0 = 0, C1 = 0;
  bool x0_half = false;                                            // sources point at fp16 data (both, when concatenated)
  int N = 0, H = 0, W = 0; int mode = 0; int silu = 1;
  const double* stats0 = nullptr; const double* stats1 = nullptr;   // per-(sample, channel) sum / sumsq of each source
  int groups = 32; float eps = 1e-5f;
  const float* gamma = nullptr; const float* beta = nullptr;        // [C0 + C1]
  const float* film = nullptr; int film_ld = 0, film_off = 0;       // optional FiLM table (scale | shift)
  bool film_add = false;                                            // table holds ONE row per channel, added before the norm
  void* out_act = nullptr; void* out_raw16 = nullptr; float* out_raw32 = nullptr;
  void* out_lo = nullptr;      // optional low half of a two-term fp16 split of the output (fp16-source same-resolution path only)
};
void launch_gn_apply(const GnApplyDesc& d, cudaStream_t s);
// the affine of the same GroupNorm as a table float2 (A, B)[N][C0 + C1] (x0 / x1 / out_* of the descriptor are not used)
void launch_gn_coeff(const GnApplyDesc& d, void* out_ab, cudaStream_t s);
// eps[n][c][h][w] = bias[c] + sum_tap Y[n][h+dy][w+dx][tap*Co + c]  (output head, see eps_gather_kernel)
void launch_eps_gather(const float* Y, const float* bias, float* eps, int N, int H, int W, int Co, int ldy, cudaStream_t s);
// plain Downsample2d / Upsample2d layers (resblock_updown=False): see elementwise.cuh
void launch_im2col_s2(const void* x16, void* col, int N, int H, int W, int C, cudaStream_t s);
void launch_upsample2x_h16(const void* x16, void* out, int N, int H, int W, int C, cudaStream_t s);
void launch_resample_f32(const float* x, float* out, void* out16, int N, int H, int W, int C, int mode, cudaStream_t s);
void launch_pack_input(const float* x, void* out, int N, int Nx, int Cin, int HW, cudaStream_t s);

struct CondPackDesc {
  const float* x = nullptr; const float* y = nullptr; const float* mask = nullptr; const float* mask_rgb = nullptr;
  const float* noise = nullptr; void* out = nullptr; int N = 0, Nx = 0, H = 0, W = 0; int kind = 0;
  uint64_t seed = 0; uint32_t stream = 0; const int* stream_dev = nullptr;
};
void launch_cond_pack(const CondPackDesc& d, cudaStream_t s);   // sampler.cu
void launch_cfg_mix(const float* eps2, float* out, size_t count, float strength, cudaStream_t s);   // sampler.cu

void launch_posenc(const int64_t* t, int Nt, const float* freqs, int half, float* out, int N, cudaStream_t s);
// FiLM table (all ResBlock emb_layers as one product): out = silu(emb) * Wp^T + bias, Wp in the swizzled K-chunk-major
// layout described at film_table_kernel; x_t is a [ceil(N/32)][K][32] fp32 scratch.
void launch_film_table(const float* emb, const float* Wp, const float* bias, float* x_t, float* out, int N, int K, int O, cudaStream_t s);
void launch_linear(const float* in, const float* W, const float* bias, float* out, int N, int K, int O, int silu_in,
                   const float* label_emb, const int64_t* classes, int Ncls, cudaStream_t s);

}  // namespace ivid
