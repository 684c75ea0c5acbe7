// Implicit-GEMM convolution (3x3 pad 1 / 1x1) for NHWC fp16 activations on tcgen05 tensor cores.
//
// Replaces, on the ADM UNet hot path, every nn.Conv2d / nn.Conv1d the reference issues through cuDNN:
//   ResBlock2d in_layers[2] / out_layers[3] / skip_connection   (reference diffusion/backbones/adm.py:160,182,190)
//   AttentionBlock qkv / proj_out                                (adm.py:275,278)
//   input conv / final out conv                                  (adm.py:369,486)
//
// GEMM view:  D[M = N*H*W pixels, Cout] = A[M, K] * B[Cout, K]^T,   K = sum over segments of taps*C_seg.
//   A is never materialised (no im2col buffer): for every filter tap the producer issues one 4-D TMA box load
//   (64 channels x TW x TH x TN pixels) at the tap-shifted coordinate; TMA's out-of-bounds zero fill implements the
//   conv zero padding.  A second "segment" lets a 1x1 skip convolution over a different tensor (raw x) accumulate
//   into the same TMEM tile as extra K slabs, so  skip(x) + conv(h)  is one kernel.
//   B (weights) is packed [Cout_pad][K] fp16, K-major, loaded by 2-D TMA.
// Both operands land in shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma consumes directly.
//
// Persistent, warp-specialised:  warp0 = TMA producer, warp1 = MMA issuer (single thread), warp2 = TMEM allocator,
// warps4-7 = epilogue (TMEM -> registers -> bias / residual / cast -> global).  Accumulators are double-buffered in TMEM
// (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "common.cuh"

namespace ivid {

struct ConvGemmParams {
  int N, H, W;                 // batch and spatial size of the conv input == output (stride 1)
  int TW, TH, TN;              // pixel tile (TW*TH*TN == 128)
  int tiles_w, tiles_h, tiles_n;
  int n_blocks;                // Cout_pad / BN
  int num_tiles;               // tiles_w*tiles_h*tiles_n*n_blocks
  int seg_chunks[2];           // channels/64 of each K segment (0 = segment unused)
  int seg_taps[2];             // 9 (3x3) or 1 (1x1)
  int Cout;                    // valid output channels
  int ldc;                     // output channel stride (elements) for NHWC modes
  int ldr;                     // residual channel stride (elements)
  int out_mode;                // 0 = fp32 NHWC, 1 = fp16 NHWC, 2 = fp32 NCHW (Cout planes)
  const float* bias;           // [Cout_pad] fp32
  const float* residual;       // fp32 NHWC or nullptr
  void* out;
  double* stats;               // optional [N][Cout][2] per-(sample, channel) sum / sum-of-squares of the output (GroupNorm)
};

template <int BN>
struct ConvGemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;                  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : (BN == 64) ? 8 : 10;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAT_BYTES = 8 * BN * 4;                                // [4 warps][sum|sumsq][BN] fp32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STAT_BYTES + 1024;   // +1024 alignment slack
  static constexpr int THREADS = 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_area = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(bar_area + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc<Cfg::TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblks = p.seg_chunks[0] * p.seg_taps[0] + p.seg_chunks[1] * p.seg_taps[1];
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nblk = tile % p.n_blocks;
      const int mt = tile / p.n_blocks;
      const int tn = mt / tiles_per_img;
      const int rem = mt - tn * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      const int n0 = tn * p.TN, h0 = th * p.TH, w0 = tw * p.TW;
      int kcol = 0;
#pragma unroll 1
      for (int seg = 0; seg < 2; ++seg) {
        const CUtensorMap* mapA = seg == 0 ? &mapA0 : &mapA1;
        const int taps = p.seg_taps[seg];
        const int chunks = p.seg_chunks[seg];
        if (chunks == 0) continue;
#pragma unroll 1
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
          const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
#pragma unroll 1
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            tma_load_4d(mapA, &full_bar[stage], sa, ch * 64, w0 + dx, h0 + dy, n0);
            tma_load_2d(&mapB, &full_bar[stage], sb, kcol, nblk * BN);
            kcol += 64;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================================== MMA issuer =====================================
    constexpr uint32_t idesc = make_idesc_f16(128, BN, false, false, false);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
#pragma unroll 1
      for (int kb = 0; kb < kblks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t da = make_smem_desc_sw128(sa, 1024, 16);
        const uint64_t db = make_smem_desc_sw128(sb, 1024, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
          mma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);   // frees the smem slot once the MMAs above have consumed it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // row of the 128-pixel tile
    const int pw = row % p.TW;
    const int ph = (row / p.TW) % p.TH;
    const int pn = row / (p.TW * p.TH);
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = (BN >= 32) ? 32 : 16;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nblk = tile % p.n_blocks;
      const int mt = tile / p.n_blocks;
      const int tn = mt / tiles_per_img;
      const int rem = mt - tn * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int tw = rem - th * p.tiles_w;
      const int n = tn * p.TN + pn, h = th * p.TH + ph, w = tw * p.TW + pw;
      const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
      const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      const bool do_stats = (CH == 32) && (p.stats != nullptr);
      const int n_warp = tn * p.TN + (quarter * 32) / (p.TW * p.TH);     // sample of this warp's 32 rows (TW*TH >= 32)
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t r[CH];
        if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr + c0, r);
        else tmem_ld_32x32b_x16(taddr + c0, r);
        tc_wait_ld();
        const int col0 = nblk * BN + c0;
        if (col0 >= p.Cout) continue;                   // padded output columns (warp-uniform)
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = valid ? __uint_as_float(r[j]) + __ldg(p.bias + col0 + j) : 0.f;
        if (valid) {
          if (p.out_mode == 2) {
            // fp32 NCHW planes (final eps output): Cout is tiny (4)
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              const int c = col0 + j;
              if (c < p.Cout) o[((static_cast<size_t>(n) * p.Cout + c) * p.H + h) * p.W + w] = v[j];
            }
          } else {
            if (p.residual != nullptr) {
              const float* rp = p.residual + pix * p.ldr + col0;
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                if (col0 + j < p.Cout) {
                  const float4 t = ldg_f4(rp + j);
                  v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                }
              }
            }
            if (p.out_mode == 0) {
              float* o = reinterpret_cast<float*>(p.out) + pix * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < CH; j += 4)
                if (col0 + j < p.Cout) stg_f4(o + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            } else {
              __half* o = reinterpret_cast<__half*>(p.out) + pix * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                if (col0 + j < p.Cout) {
                  uint4 pk;
                  pk.x = pack_h2(v[j], v[j + 1]);
                  pk.y = pack_h2(v[j + 2], v[j + 3]);
                  pk.z = pack_h2(v[j + 4], v[j + 5]);
                  pk.w = pack_h2(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(o + j) = pk;
                }
              }
            }
          }
        }
        if constexpr (CH == 32) {
          if (do_stats) {
            // GroupNorm statistics of the tensor being written (consumed by the NEXT norm): per-column sum and sum of
            // squares over this warp's 32 rows by a transposing butterfly (31 shuffles per quantity); lane l ends up
            // with column c0 + l.
            float q[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) q[j] = v[j] * v[j];
#pragma unroll
            for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < cnt / 2; ++i) {
                const float send_v = up ? v[i] : v[i + cnt / 2];
                const float keep_v = up ? v[i + cnt / 2] : v[i];
                v[i] = keep_v + __shfl_xor_sync(0xffffffffu, send_v, off);
                const float send_q = up ? q[i] : q[i + cnt / 2];
                const float keep_q = up ? q[i + cnt / 2] : q[i];
                q[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
              }
            }
            if (p.TN == 1) {
              stat_smem[(quarter * 2 + 0) * BN + c0 + lane] = v[0];
              stat_smem[(quarter * 2 + 1) * BN + c0 + lane] = q[0];
            } else if (n_warp < p.N && col0 + lane < p.Cout) {
              double* st = p.stats + (static_cast<size_t>(n_warp) * p.Cout + col0 + lane) * 2;
              atomicAdd(st, static_cast<double>(v[0]));
              atomicAdd(st + 1, static_cast<double>(q[0]));
            }
          }
        }
      }
      if (do_stats && p.TN == 1) {
        // combine the four epilogue warps (same sample when TN == 1), then one double atomic per (column, moment)
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        const int t = threadIdx.x - 128;
        for (int c = t; c < BN; c += 128) {
          const int col = nblk * BN + c;
          if (col < p.Cout && n_warp < p.N) {
            float ssum = 0.f, qsum = 0.f;
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) { ssum += stat_smem[(w4 * 2 + 0) * BN + c]; qsum += stat_smem[(w4 * 2 + 1) * BN + c]; }
            double* st = p.stats + (static_cast<size_t>(tn) * p.Cout + col) * 2;
            atomicAdd(st, static_cast<double>(ssum));
            atomicAdd(st + 1, static_cast<double>(qsum));
          }
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace ivid
