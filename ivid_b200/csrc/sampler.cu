// DDPM / DDIM samplers: float64 schedule tables (host), device coefficient table, fused step kernels, whole-loop driver.
#include "sampler.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "sampler.cuh"

namespace ivid {

// --------------------------------------------------------------------------------------------------
// tiny state kernels: keep the per-step scalars on the device so that a step never needs an H2D copy
// --------------------------------------------------------------------------------------------------
struct StepState {
  int t_index;       // coefficient-table row of the model timestep
  int t_prev;        // DDIM: previous actual step
  int stream;        // Philox stream (step counter)
  int pad;
};

__global__ void set_step_kernel(StepState* st, int64_t* t_model, int N, int t_index, int t_prev, int stream) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->t_index = t_index;
    st->t_prev = t_prev;
    st->stream = stream;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) t_model[i] = t_index;
}

// same, with the step read from the caller's device tensors (sample_once(x_t, t[, t_prev]) of the reference passes [N]
// tensors; reading element 0 here removes the device->host sync an int(t[0]) would cost).  Out-of-range steps are
// clamped into the table (the host path raises instead).
__global__ void set_step_dev_kernel(StepState* st, int64_t* t_model, int N, const int64_t* t_dev, const int64_t* t_prev_dev,
                                    int ddim, int T) {
  long long t = t_dev[0];
  long long ti = ddim ? t - 1 : t;
  ti = ti < 0 ? 0 : (ti > T - 1 ? T - 1 : ti);
  long long tp = (ddim && t_prev_dev != nullptr) ? t_prev_dev[0] : 0;
  tp = tp < 0 ? 0 : (tp > T ? T : tp);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->t_index = static_cast<int>(ti);
    st->t_prev = static_cast<int>(tp);
    st->stream = static_cast<int>(t);
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) t_model[i] = ti;
}

__global__ void fill_classes_kernel(const int64_t* classes, int64_t* out, int N) {
  // [classes..., -1 ...]: conditional half then null-class half (classifier_free_guidance.py:39-42)
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < 2 * N; i += blockDim.x * gridDim.x)
    out[i] = i < N ? classes[i] : -1;
}

void launch_cfg_mix(const float* eps2, float* out, size_t count, float strength, cudaStream_t s) {
  IVID_REQUIRE(count % 4 == 0, "cfg mix: element count must be a multiple of 4");
  const size_t n4 = count / 4;
  const int grid = static_cast<int>(std::min<size_t>((n4 + 255) / 256, static_cast<size_t>(sm_count()) * 8));
  cfg_mix_kernel<<<std::max(grid, 1), 256, 0, s>>>(eps2, out, n4, strength);
  IVID_CHECK_CUDA(cudaGetLastError());
}

void launch_cond_pack(const CondPackDesc& d, cudaStream_t s) {
  CondPackParams cp;
  cp.x = d.x; cp.y = d.y; cp.mask = d.mask; cp.mask_rgb = d.mask_rgb; cp.noise = d.noise;
  cp.out = reinterpret_cast<__half*>(d.out); cp.N = d.N; cp.Nx = d.Nx; cp.H = d.H; cp.W = d.W; cp.kind = d.kind;
  cp.seed = d.seed; cp.stream = d.stream; cp.stream_dev = d.stream_dev;
  IVID_REQUIRE(d.kind == 1 || d.kind == 2, "cond inputs: kind must be 1 (inpaint) or 2 (super-resolution)");
  IVID_REQUIRE(d.y != nullptr, "cond inputs: y is required");
  IVID_REQUIRE(d.kind != 1 || d.mask != nullptr, "cond inputs: mask is required for inpainting");
  const size_t items = static_cast<size_t>(d.N) * d.H * d.W;
  const int grid = static_cast<int>(std::min<size_t>((items + 255) / 256, static_cast<size_t>(sm_count()) * 16));
  cond_pack_kernel<<<std::max(grid, 1), 256, 0, s>>>(cp);
  IVID_CHECK_CUDA(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------------
// Sampler
// --------------------------------------------------------------------------------------------------
Sampler::Sampler(const double* betas, int T) : T_(T) {
  IVID_REQUIRE(T >= 1, "timesteps must be positive");
  betas_.assign(betas, betas + T);
  for (double b : betas_) IVID_REQUIRE(b > 0.0 && b <= 1.0, "betas must be in (0, 1]");   // gaussian_diffusion.py:38
  // ddpm.py:26-41 / ddim.py:26-31, all float64
  acp_.resize(T); acp_prev_.resize(T); srac_.resize(T); srm1_.resize(T); pvar_.resize(T); plogvar_.resize(T);
  pc1_.resize(T); pc2_.resize(T);
  double prod = 1.0;
  for (int i = 0; i < T; ++i) {
    prod *= (1.0 - betas_[i]);
    acp_[i] = prod;
  }
  for (int i = 0; i < T; ++i) acp_prev_[i] = i == 0 ? 1.0 : acp_[i - 1];
  for (int i = 0; i < T; ++i) {
    srac_[i] = std::sqrt(1.0 / acp_[i]);
    srm1_[i] = std::sqrt(1.0 / acp_[i] - 1.0);
    pvar_[i] = betas_[i] * (1.0 - acp_prev_[i]) / (1.0 - acp_[i]);
    pc1_[i] = betas_[i] * std::sqrt(acp_prev_[i]) / (1.0 - acp_[i]);
    pc2_[i] = (1.0 - acp_prev_[i]) * std::sqrt(1.0 - betas_[i]) / (1.0 - acp_[i]);
  }
  for (int i = 0; i < T; ++i) plogvar_[i] = std::log(pvar_[(i == 0 && T > 1) ? 1 : i]);
}

Sampler::~Sampler() {
  if (d_table_) cudaFree(d_table_);
  if (d_state_) cudaFree(d_state_);
  if (d_t_) cudaFree(d_t_);
  if (d_classes2_) cudaFree(d_classes2_);
  if (d_eps_) cudaFree(d_eps_);
  if (d_xtmp_) cudaFree(d_xtmp_);
}

const std::vector<double>& Sampler::table(int which) const {
  switch (which) {
    case 0: return acp_;
    case 1: return acp_prev_;
    case 2: return srac_;
    case 3: return srm1_;
    case 4: return pvar_;
    case 5: return plogvar_;
    case 6: return pc1_;
    case 7: return pc2_;
    default: throw Error(kErrInvalidArgument, "unknown table id");
  }
}

void Sampler::ensure_device(int N2, size_t eps_elems) {
  if (!d_table_) {
    std::vector<StepCoef> rows(T_);
    for (int i = 0; i < T_; ++i) {
      rows[i].sqrt_recip_acp = static_cast<float>(srac_[i]);
      rows[i].sqrt_recipm1_acp = static_cast<float>(srm1_[i]);
      rows[i].post_mean_coef1 = static_cast<float>(pc1_[i]);
      rows[i].post_mean_coef2 = static_cast<float>(pc2_[i]);
      rows[i].post_logvar = static_cast<float>(plogvar_[i]);
      rows[i].acp = static_cast<float>(acp_[i]);
      rows[i].acp_prev = static_cast<float>(acp_prev_[i]);
      rows[i].pad = 0.f;
    }
    IVID_CHECK_CUDA(cudaMalloc(&d_table_, sizeof(StepCoef) * T_));
    IVID_CHECK_CUDA(cudaMemcpy(d_table_, rows.data(), sizeof(StepCoef) * T_, cudaMemcpyHostToDevice));
    IVID_CHECK_CUDA(cudaMalloc(&d_state_, sizeof(StepState)));
  }
  if (N2 > cap_n_) {
    if (d_t_) cudaFree(d_t_);
    if (d_classes2_) cudaFree(d_classes2_);
    IVID_CHECK_CUDA(cudaMalloc(&d_t_, sizeof(int64_t) * N2));
    IVID_CHECK_CUDA(cudaMalloc(&d_classes2_, sizeof(int64_t) * N2));
    cap_n_ = N2;
  }
  if (eps_elems > cap_eps_) {
    if (d_eps_) cudaFree(d_eps_);
    if (d_xtmp_) cudaFree(d_xtmp_);
    IVID_CHECK_CUDA(cudaMalloc(&d_eps_, eps_elems * 4));
    IVID_CHECK_CUDA(cudaMalloc(&d_xtmp_, eps_elems * 4));
    cap_eps_ = eps_elems;
  }
}

void Sampler::step(Unet& unet, const float* x_t, float* x_prev, float* pred_x0, int N, int t, int t_prev,
                   const ivid_step_args_t& a, int stream_id, cudaStream_t stream, const int64_t* t_dev,
                   const int64_t* t_prev_dev) {
  const UnetConfig& uc = unet.cfg();
  const int C = uc.out_channels, S = uc.image_size, HW = S * S;
  IVID_REQUIRE(N >= 1, "batch must be positive");
  IVID_REQUIRE(HW % 4 == 0, "image size");
  const bool ddim = a.kind == 1;
  IVID_REQUIRE(a.kind == 0 || a.kind == 1, "sampler kind must be 0 (DDPM) or 1 (DDIM)");
  const int t_index = ddim ? t - 1 : t;     // ddim.py:81 calls the model with t - 1
  IVID_REQUIRE(t_dev != nullptr || (t_index >= 0 && t_index < T_), "t out of range");
  IVID_REQUIRE(t_dev != nullptr || !ddim || (t_prev >= 0 && t_prev <= T_), "t_prev out of range");
  // classifier-free guidance: one batch-2N forward when strength > 0 and the model is class conditional
  const bool has_classes = a.classes_dev != nullptr;
  const bool cfg_two = a.use_cfg && has_classes && a.strength > 0.0f;
  // inpaint_cfg.py:77-78 / sr_cfg.py:53-54: classes None -> single null-class forward, no (1+s) scaling
  // strength < 0: (1 + strength) * eps of ONE forward (classifier_free_guidance.py:40-41); the conditional frameworks skip even
  // that when classes is None
  const bool scale_only = a.use_cfg && a.strength < 0.0f && (has_classes || a.cond.kind == 0);
  const int Nf = cfg_two ? 2 * N : N;
  IVID_CHECK_CUDA(cudaSetDevice(unet.device()));      // before any allocation: a direct C-ABI caller may be on another device
  ensure_device(Nf, static_cast<size_t>(Nf) * C * HW);

  if (t_dev != nullptr)
    set_step_dev_kernel<<<1, 128, 0, stream>>>(reinterpret_cast<StepState*>(d_state_), d_t_, Nf, t_dev, t_prev_dev, ddim ? 1 : 0, T_);
  else
    set_step_kernel<<<1, 128, 0, stream>>>(reinterpret_cast<StepState*>(d_state_), d_t_, Nf, t_index, t_prev, stream_id);
  IVID_CHECK_CUDA(cudaGetLastError());
  const int64_t* cls = nullptr;
  if (has_classes) {
    if (cfg_two) {
      // [classes, -1 ...] of the batch-2N forward; inside run() it is filled once for the whole reverse process
      if (!(classes2_ready_ && classes2_src_ == a.classes_dev && classes2_n_ == N)) {
        fill_classes_kernel<<<1, 256, 0, stream>>>(a.classes_dev, d_classes2_, N);
        IVID_CHECK_CUDA(cudaGetLastError());
      }
      cls = d_classes2_;
    } else {
      cls = a.classes_dev;
    }
  }
  ivid_cond_t cond = a.cond;
  // in-kernel noise of the conditional inputs: Philox(seed', step) with the step read from the device-resident step state
  // (not passed by value: consecutive steps then replay the same CUDA graph of the forward)
  if (cond.kind != 0 && cond.noise_dev == nullptr) { cond.seed = a.seed ^ 0x9E3779B97F4A7C15ull; cond.stream_id = 0; }

  StepParams p;
  std::memset(&p, 0, sizeof(p));         // padding bytes are part of the fused route's graph key
  p.x_t = x_t; p.eps = d_eps_; p.noise = a.step_noise_dev; p.x_prev = x_prev; p.pred_x0 = pred_x0;
  p.table = reinterpret_cast<const StepCoef*>(d_table_);
  p.t_index = &reinterpret_cast<StepState*>(d_state_)->t_index;
  p.t_prev = &reinterpret_cast<StepState*>(d_state_)->t_prev;
  p.N = N; p.C = C; p.HW = HW;
  // strength <= 0 with classes: (1 + strength) * eps_c without the null-class forward (classifier_free_guidance.py:40-41)
  p.cfg = cfg_two ? 1 : (scale_only ? 2 : 0);
  p.strength = a.strength;
  p.clip = a.clip_denoised; p.eta = a.eta; p.seed = a.seed; p.stream = 0;
  p.stream_dev = &reinterpret_cast<StepState*>(d_state_)->stream;
  GuideParams& g = p.g;
  g.rgb = a.replace_rgb_dev; g.rgb_mask = a.replace_rgb_mask_dev;
  g.depth = a.replace_depth_dev; g.depth_mask = a.replace_depth_mask_dev; g.convex = a.constrain_depth_dev;
  g.w_rgb = static_cast<float>(a.replace_rgb_weight); g.w_rgb_c = static_cast<float>(1.0 - a.replace_rgb_weight);
  g.w_depth = static_cast<float>(a.replace_depth_weight); g.w_depth_c = static_cast<float>(1.0 - a.replace_depth_weight);
  g.w_convex = static_cast<float>(a.constrain_depth_weight); g.w_convex_c = static_cast<float>(1.0 - a.constrain_depth_weight);
  IVID_REQUIRE(g.rgb == nullptr || g.rgb_mask != nullptr, "replace_rgb needs its mask");
  IVID_REQUIRE(g.depth == nullptr || g.depth_mask != nullptr, "replace_depth needs its mask");
  IVID_REQUIRE(g.convex == nullptr || g.depth != nullptr, "constrain_depth is applied inside replace_depth (ddim.py:90-95)");
  IVID_REQUIRE(!ddim ? (g.rgb == nullptr && g.depth == nullptr) : true, "replace/constrain guidance is DDIM-only");

  unet.set_cond_stream_dev(cond.kind != 0 && cond.noise_dev == nullptr ? &reinterpret_cast<StepState*>(d_state_)->stream : nullptr);
  // Fused route: the output head's last kernel IS the step (head_step_kernel): eps never reaches HBM and the update is the last
  // node of the forward's CUDA graph.  Not taken when per-step pointers change every step (injected noise / trajectories inside
  // run(): every step would need its own graph) or when the model has no tap-column head.
  static const bool fuse_ok = getenv("IVID_NO_FUSED_STEP") == nullptr;
  const bool fuse = fuse_ok && !no_fuse_ && unet.can_fuse_head() && C == 4 && S % 4 == 0;
  if (fuse) {
    p.eps = nullptr;
    HeadStepParams hp;
    std::memset(&hp, 0, sizeof(hp));
    hp.sp = p; hp.H = S; hp.W = S;
    HeadHook hook;
    uint64_t h = 1469598103934665603ull ^ (ddim ? 0x9E37ull : 0ull);         // FNV-1a over everything the launcher bakes in
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(&p);
    for (size_t i = 0; i < sizeof(StepParams); ++i) { h ^= bytes[i]; h *= 1099511628211ull; }
    hook.key = h | 1ull;
    hook.launch = [hp, ddim](const float* Y, const float* bias, int, int H, int W, int Co, int ldy, cudaStream_t st) mutable {
      IVID_REQUIRE(Co == 4 && W % 4 == 0, "fused head step: 4 output channels, width % 4 == 0");
      HeadStepParams q = hp;
      q.Y = Y; q.bias = bias; q.H = H; q.W = W; q.ldy = ldy;
      const size_t groups = static_cast<size_t>(q.sp.N) * H * (W / 4);
      const int grid = static_cast<int>(std::min<size_t>((groups + 255) / 256, static_cast<size_t>(sm_count()) * 8));
      if (ddim) head_step_kernel<true><<<std::max(grid, 1), 256, 0, st>>>(q);
      else head_step_kernel<false><<<std::max(grid, 1), 256, 0, st>>>(q);
      IVID_CHECK_CUDA(cudaGetLastError());
    };
    unet.forward(x_t, N, cond.kind ? &cond : nullptr, d_t_, cls, nullptr, Nf, stream, &hook);
    unet.set_cond_stream_dev(nullptr);
    return;
  }
  unet.forward(x_t, N, cond.kind ? &cond : nullptr, d_t_, cls, d_eps_, Nf, stream);
  unet.set_cond_stream_dev(nullptr);
  const size_t total4 = static_cast<size_t>(N) * C * HW / 4;
  const int grid = static_cast<int>(std::min<size_t>((total4 + 255) / 256, static_cast<size_t>(sm_count()) * 8));
  if (ddim) step_kernel<true><<<std::max(grid, 1), 256, 0, stream>>>(p);
  else step_kernel<false><<<std::max(grid, 1), 256, 0, stream>>>(p);
  IVID_CHECK_CUDA(cudaGetLastError());
}

void Sampler::run(Unet& unet, float* x, int N, int steps, const ivid_step_args_t& a, const float* noise_all,
                  const float* cond_noise_all, float* traj_x0, float* traj_xt, cudaStream_t stream) {
  const UnetConfig& uc = unet.cfg();
  const size_t img = static_cast<size_t>(N) * uc.out_channels * uc.image_size * uc.image_size;
  const bool ddim = a.kind == 1;
  if (!ddim) steps = T_;
  IVID_REQUIRE(steps >= 1 && steps <= T_, "steps out of range");
  const int jump = T_ / steps;                     // ddim.py:153
  IVID_CHECK_CUDA(cudaSetDevice(unet.device()));
  ensure_device(2 * N, 2 * img);
  float* bufs[2] = {x, d_xtmp_};                   // ping-pong; the result is copied back to x if it ends in d_xtmp_
  int cur = 0;
  // per denoising step the host then issues three calls: the step-state kernel, ONE CUDA-graph launch (the whole batch-2N
  // forward) and the fused guidance-mix + x_{t-1} update
  struct Ready { Sampler* s; ~Ready() { s->classes2_ready_ = false; s->no_fuse_ = false; } } ready_guard{this};
  no_fuse_ = noise_all != nullptr || cond_noise_all != nullptr || traj_x0 != nullptr || traj_xt != nullptr;
  if (a.use_cfg && a.classes_dev != nullptr && a.strength > 0.0f) {
    fill_classes_kernel<<<1, 256, 0, stream>>>(a.classes_dev, d_classes2_, N);
    IVID_CHECK_CUDA(cudaGetLastError());
    classes2_ready_ = true; classes2_src_ = a.classes_dev; classes2_n_ = N;
  }
  for (int i = 0; i < steps; ++i) {
    int t, t_prev;
    if (ddim) { t = jump * (steps - i); t_prev = jump * (steps - 1 - i); }   // ddim.py:154
    else { t = T_ - 1 - i; t_prev = 0; }                                      // ddpm.py:177
    ivid_step_args_t ai = a;
    ai.step_noise_dev = noise_all ? noise_all + static_cast<size_t>(i) * img : nullptr;
    if (cond_noise_all && ai.cond.kind == 1)
      ai.cond.noise_dev = cond_noise_all + static_cast<size_t>(i) * N * 4 * uc.image_size * uc.image_size;
    float* dst = traj_xt ? traj_xt + static_cast<size_t>(i) * img : bufs[cur ^ 1];
    float* x0 = traj_x0 ? traj_x0 + static_cast<size_t>(i) * img : nullptr;
    const float* src = (traj_xt && i > 0) ? traj_xt + static_cast<size_t>(i - 1) * img : bufs[cur];
    if (traj_xt && i == 0) src = x;
    step(unet, src, dst, x0, N, t, t_prev, ai, i, stream);
    if (!traj_xt) cur ^= 1;
  }
  const float* last = traj_xt ? traj_xt + static_cast<size_t>(steps - 1) * img : bufs[cur];
  if (last != x) IVID_CHECK_CUDA(cudaMemcpyAsync(x, last, img * 4, cudaMemcpyDeviceToDevice, stream));
}

}  // namespace ivid
