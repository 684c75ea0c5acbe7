/*
 * ivid_b200 — C ABI of the B200-native (sm_100a) multiview RGBD diffusion sampling hot path.
 *
 * The reference (JeffreyXiang/ivid) has no FFI layer: its boundary for this path is the Python class surface that
 * inference/sample.py resolves by name (sample.py:183-184,191-192,47-50).  The Python package `ivid_b200` mirrors
 * those classes and binds the entry points below through ctypes.  Each entry point cites the
 * reference interface it replaces.  All pointers are plain host or device pointers, sizes are explicit, no torch
 * types cross this boundary.  Every function returns 0 on success or one of the IVID_ERR_* codes; the message is
 * available (thread-local) from ivid_last_error().  The library never aborts.
 *
 * Unless stated otherwise `*_dev` pointers are device pointers on the handle's device, tensors are dense fp32 NCHW
 * (the layout of the reference's torch tensors), and work is enqueued on `stream` (a cudaStream_t passed as void*).
 */
#ifndef IVID_B200_H_
#define IVID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVID_OK 0
#define IVID_ERR_INVALID_ARGUMENT 1 /* reference: Python `assert` (adm.py:540-549, ddpm.py:86)      -> AssertionError      */
#define IVID_ERR_NOT_IMPLEMENTED 2  /* reference: NotImplementedError (frameworks/utils.py:37)      -> NotImplementedError */
#define IVID_ERR_CUDA 3             /* CUDA runtime / driver failure                                  -> RuntimeError        */
#define IVID_ERR_STATE 4            /* call order violation (e.g. forward before finalize)            -> RuntimeError        */

typedef struct ivid_unet ivid_unet_t;
typedef struct ivid_sampler ivid_sampler_t;

const char* ivid_last_error(void);
int ivid_version(void);
/* Number of SMs / compute capability of `device` (diagnostics; fails without a GPU). */
int ivid_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------------------------
 * ADM UNet backbone  — replaces diffusion.backbones.AdmUnet2d (reference diffusion/backbones/adm.py:289-566)
 * ------------------------------------------------------------------------------------------------------------------ */

/* AdmUnet2d.__init__ (adm.py:318-337).  `cfg_json` is the "backbone.args" object of a reference config file
 * (configs/(name).json), passed through unchanged.  No GPU work happens here (usable on a CPU-only host). */
int ivid_unet_create(const char* cfg_json, ivid_unet_t** out);
int ivid_unet_destroy(ivid_unet_t* h);

/* state_dict schema (adm.py:357-366,367-487; the 494-key contract of SURVEY.md §8b): enumerate names/shapes. */
int ivid_unet_num_params(const ivid_unet_t* h, int* count);
int ivid_unet_param_info(const ivid_unet_t* h, int index, const char** name, int64_t shape[4], int* ndim,
                         int* is_buffer);
/* load_state_dict (sample.py:187): copy one fp32 host tensor in reference layout ([Cout,Cin,kh,kw], [O,I], ...). */
int ivid_unet_set_param(ivid_unet_t* h, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* .cuda() : pack all parameters (fp16 K-major GEMM operands, fp32 norms/embeddings) into one device arena. */
int ivid_unet_finalize(ivid_unet_t* h, int device);
/* Device arena (for the NCCL weight broadcast at init, sample.py:186-195 loads per rank instead). */
int ivid_unet_weight_arena(const ivid_unet_t* h, void** dev_ptr, uint64_t* bytes);

/* AdmUnet2d.forward(x, times, classes) (adm.py:526-566).
 *   x_dev       fp32 [Nx, in_channels, H, W]; sample n of the batch reads x[n % Nx] (Nx == N for the plain call)
 *   t_dev       int64 [N]  (diffusion step minus 1, as the reference passes it)
 *   classes_dev int64 [N] or NULL; -1 selects the null class (adm.py:551-553)
 *   eps_dev     fp32 [N, out_channels, H, W] */
int ivid_unet_forward(ivid_unet_t* h, const float* x_dev, int Nx, const int64_t* t_dev, const int64_t* classes_dev,
                      float* eps_dev, int N, void* stream);

/* Parity aid (per-layer taps, tests/test_gpu_unet.py): output of the module `layer` (reference module path such as
 * "input_blocks.3.0", "middle_block.1", "output_blocks.14.0"; the stem is "input_blocks.0.0"; "emb" = time + class embedding
 * [N, 4*model_channels, 1, 1]; "film" = the stacked emb_layers outputs of all ResBlocks) of the LAST forward of batch
 * N, as fp32 NCHW on the host.  host_out == NULL only queries the shape.  Synchronises the device. */
int ivid_unet_debug_tap(ivid_unet_t* h, int N, const char* layer, float* host_out, uint64_t capacity, int* C, int* H, int* W);

/* Profiling aid (bench.py roofline): between begin/end every kernel launch of ivid_unet_forward is bracketed by CUDA
 * events on the launching stream; end returns JSON {"kernel family": {"launches","ms","flops","bytes"}} where flops /
 * bytes are the ALGORITHMIC figures of DESIGN.md.  Forwards issued while profiling synchronise the stream. */
int ivid_unet_profile_begin(ivid_unet_t* h);
int ivid_unet_profile_end(ivid_unet_t* h, char* json_out, int capacity);

/* Conditional inputs assembled on the fly (never materialised in fp32):
 *   kind 1: InpaintCFG.make_cond_inputs (frameworks/inpaint_cfg.py:24-49): cat[x, mask_rgb, y_rgb*m_rgb+z*(1-m_rgb),
 *           y_d*m+z*(1-m), m];  noise_dev = injected z [Nx,4,H,W] or NULL (in-kernel Philox(seed, stream)).
 *   kind 2: SuperResCFG.make_cond_inputs (frameworks/sr_cfg.py:23-36): cat[x, bilinear_up2(y)], y is [Nx,4,H/2,W/2]. */
typedef struct {
  int kind;              /* 0 none, 1 inpaint, 2 super-resolution */
  const float* y_dev;
  const float* mask_dev;
  const float* mask_rgb_dev;
  const float* noise_dev;
  uint64_t seed;
  uint32_t stream_id;
} ivid_cond_t;
int ivid_unet_forward_cond(ivid_unet_t* h, const float* x_dev, int Nx, const ivid_cond_t* cond, const int64_t* t_dev,
                           const int64_t* classes_dev, float* eps_dev, int N, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Samplers — replace diffusion.samplers.DdpmSampler / DdimSampler (samplers/ddpm.py:12-187, samplers/ddim.py:12-165)
 * together with the framework's model_inference (classifier_free_guidance.py:23-42, inpaint_cfg.py:61-83,
 * sr_cfg.py:39-60).
 * ------------------------------------------------------------------------------------------------------------------ */

/* DdpmSampler/DdimSampler.__init__ (ddpm.py:20-41, ddim.py:19-31): derive the float64 tables from framework.betas. */
int ivid_sampler_create(const double* betas, int timesteps, ivid_sampler_t** out);
int ivid_sampler_destroy(ivid_sampler_t* s);
/* Known-answer access to the float64 tables (which: 0 alphas_cumprod, 1 alphas_cumprod_prev, 2 sqrt_recip_acp,
 * 3 sqrt_recipm1_acp, 4 posterior_variance, 5 posterior_log_variance_clipped, 6 posterior_mean_coef1, 7 coef2). */
int ivid_sampler_table(const ivid_sampler_t* s, int which, double* out, int count);

typedef struct {
  int kind;                 /* 0 = DDPM ancestral (ddpm.py:111-131), 1 = DDIM (ddim.py:48-103) */
  int use_cfg;              /* 1: (1+strength)*eps(c) - strength*eps(null), both halves in ONE batch-2N forward */
  float strength;           /* <= 0: ONE forward, eps scaled by (1+strength) when classes are given (classifier_free_guidance.py:40-41) */
  int clip_denoised;
  float eta;
  const int64_t* classes_dev;   /* [N] or NULL */
  ivid_cond_t cond;             /* conditional-model inputs (kind 0 for the unconditional model) */
  /* multiview guidance of DdimSampler.sample_once (ddim.py:86-95); NULL pointers disable a term */
  const float* replace_rgb_dev;        /* [N,3,H,W] */
  const float* replace_rgb_mask_dev;   /* [N,1,H,W] */
  double replace_rgb_weight;
  const float* replace_depth_dev;      /* [N,1,H,W] */
  const float* replace_depth_mask_dev; /* [N,1,H,W] */
  double replace_depth_weight;
  const float* constrain_depth_dev;    /* [N,1,H,W] convex hull depth */
  double constrain_depth_weight;
  /* RNG: injected noise (parity tests) or in-kernel Philox4x32-10 keyed by (seed, step) */
  const float* step_noise_dev;         /* [N,C,H,W] noise of THIS step (ivid_sampler_step) or NULL */
  uint64_t seed;
} ivid_step_args_t;

/* sample_once: x_prev = f(x_t, t[, t_prev]).  `t` follows the reference's convention of each sampler:
 * DDPM: t in [0,T) is the step minus 1 (ddpm.py:118); DDIM: t in [1,T] actual step, t_prev in [0,T) (ddim.py:66-67).
 * pred_x0_dev may be NULL. */
int ivid_sampler_step(ivid_sampler_t* s, ivid_unet_t* unet, const float* x_t_dev, float* x_prev_dev,
                      float* pred_x0_dev, int N, int t, int t_prev, const ivid_step_args_t* args, void* stream);

/* Same step with t / t_prev read on the device from element 0 of the caller's [N] int64 tensors (the tensors
 * sample_once receives, ddpm.py:111, ddim.py:48): no device->host synchronisation.  Steps outside the schedule are
 * clamped (the host-int entry point above raises instead).  Philox stream = t. */
int ivid_sampler_step_dev(ivid_sampler_t* s, ivid_unet_t* unet, const float* x_t_dev, float* x_prev_dev,
                          float* pred_x0_dev, int N, const int64_t* t_dev, const int64_t* t_prev_dev,
                          const ivid_step_args_t* args, void* stream);

/* ClassifierFreeGuidance.model_inference's mix alone (classifier_free_guidance.py:42): out = (1+s)*eps[0:count) -
 * s*eps[count:2*count) for the batch-2N forward's output (count = N*C*H*W, multiple of 4). */
int ivid_cfg_mix(const float* eps2n_dev, float strength, float* out_dev, uint64_t count, void* stream);

/* sample: the whole reverse process on device (ddpm.py:134-187, ddim.py:106-165); x_inout_dev holds x_T on entry
 * and the samples on return.  `steps` = DDIM step count (ignored for DDPM, which runs all T).  Optional
 * noise_all_dev [steps][N,C,H,W] / cond_noise_all_dev [steps][N,4,H,W] inject the per-step draws; traj_x0_dev /
 * traj_xt_dev ([steps][N,C,H,W]) receive pred_x_0 / pred_x_t of every step when non-NULL (the reference always keeps
 * them: ddpm.py:183-184). */
int ivid_sampler_run(ivid_sampler_t* s, ivid_unet_t* unet, float* x_inout_dev, int N, int steps,
                     const ivid_step_args_t* args, const float* noise_all_dev, const float* cond_noise_all_dev,
                     float* traj_x0_dev, float* traj_xt_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * RGBD novel-view warp — replaces rgbd_3d.utils.{linearize_depth, depth_to_mesh, aggregate_conditions, project_depth,
 * depth_edge} (rgbd_3d/utils.py:38-67,144-260,311-332,420-477) and rgbd_3d.AggregationRenderer with its GLSL shaders
 * (rgbd_3d/moderngl_renderer.py:151-340, shaders/aggregation.{vsh,fsh,csh}, clear.csh).  All source views of a batch of
 * samples stay resident on the device; modelview matrices are float32[16] row-major in mathematical orientation
 * (p_cam = M * p_world), what glm.lookAt produces (inference/sample.py:304-336).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ivid_warp ivid_warp_t;
typedef struct {
  double fov_deg;   /* sample.py:258 --fov 45 */
  double near;      /* sample.py:259 --near 0.6  (z-buffer depth <-> linear depth) */
  double far;       /* sample.py:260 --far 5 */
  double atol;      /* sample.py:261; negative = Python's None.  depth_to_mesh (utils.py:227-229): both None -> no discontinuity test, */
  double rtol;      /* sample.py:262;   exactly one None -> that tolerance is 0.  Other entry points take a negative value as 0.      */
  int erode_rgb;    /* sample.py:263 */
  double padding;   /* depth_to_mesh padding: 0 = 'frustum' (sample.py:131: ring pushed out one pixel and pulled to z = -0.1);
                       > 0 = that many pixels, ring not pulled (inference/utils.py:107 load_scene uses 32 for free-view rendering,
                       datasets/base.py:238 uses image_size for the training-pair warp); < 0 = None: no ring, n*n vertices */
} ivid_warp_params_t;

/* [AggregationRenderer(render_size, image_size, near, far) for _ in range(batch)]  (sample.py:50) */
int ivid_warp_create(int image_size, int render_size, int max_views, int batch, double near, double far, int device,
                     ivid_warp_t** out);
int ivid_warp_destroy(ivid_warp_t* w);
int ivid_warp_reset(ivid_warp_t* w);                       /* start a new batch of samples: forget all source views */
int ivid_warp_num_views(const ivid_warp_t* w, int* n);
/* For every sample of the batch: colors.append(rgb); meshes.append(depth_to_mesh(linearize_depth(depth, near, far),
 * padding='frustum', fov, modelview, atol, rtol, erode_rgb, cal_normal=True))   (sample.py:83,126-139).
 * rgbd_dev: fp32 [batch,4,H,W] sampler output in [-1,1]; modelviews_host: [batch][16] (or one shared matrix). */
int ivid_warp_add_view(ivid_warp_t* w, const float* rgbd_dev, const float* modelviews_host, int shared_modelview,
                       const ivid_warp_params_t* params, void* stream);
/* rgbd_3d.utils.depth_to_mesh(depth, padding='frustum' | pixels (params->padding), cal_normal=True, ...) with numpy in/out
 * (utils.py:144-260; the numeric padding is what inference/utils.py:107 load_scene uses for free-view rendering):
 * lin_depth_host [H][W] float32 linearised depth -> vertex buffer [(H+2)^2][9] and faces [2*(H+1)^2][3] on the host. */
int ivid_warp_mesh_from_depth(ivid_warp_t* w, const float* lin_depth_host, const float* modelview_host,
                              const ivid_warp_params_t* params, float* verts_host, uint32_t* faces_host, void* stream);
/* External meshes (numpy-facing mirror of AggregationRenderer.render, tests): vertex buffer [(H+2)^2][9] float32 =
 * position, normal, uv, flag (moderngl_renderer.py:284-289), faces [2*(H+1)^2][3] uint32, colour texture [H][W][3]. */
int ivid_warp_set_mesh(ivid_warp_t* w, int sample, int view, const float* verts_host, const uint32_t* faces_host,
                       const float* color_host, const float* modelview_host);
int ivid_warp_get_mesh(ivid_warp_t* w, int sample, int view, float* verts_host, uint32_t* faces_host, float* color_host);
/* AggregationRenderer.render(meshes, colors, modelview, fov, is_autoregressive=True) for one target view per sample.
 * Device outputs at render_size S (any may be NULL): color [batch,S,S,3], depth [batch,S,S], masks [batch,S,S] (0/1). */
int ivid_warp_render(ivid_warp_t* w, const float* target_mv_host, int shared_modelview, double fov_deg, float* color_dev,
                     float* depth_dev, float* mask_color_dev, float* mask_depth_dev, void* stream);
/* aggregate_conditions(...) (utils.py:420-477): cond_dev fp32 [batch,7,H,W] = color(3), depth, mask, mask_rgb,
 * depth_convex, all in [0,1] exactly like the numpy dict the reference returns. */
int ivid_warp_aggregate(ivid_warp_t* w, const float* target_mv_host, int shared_modelview, const ivid_warp_params_t* params,
                        float* cond_dev, void* stream);
/* The post-filter half of aggregate_conditions alone (LANCZOS on 8-bit colour, depth point sample + project_depth,
 * 7-of-9 votes, depth_edge, erosion) on caller-provided raw renders. */
int ivid_warp_postfilter(ivid_warp_t* w, const float* color_dev, const float* depth_dev, const float* mask_color_dev,
                         const float* mask_depth_dev, const ivid_warp_params_t* params, float* cond_dev, void* stream);

/* Free-view frame resolve (inference/render.py:74-84) of the LAST ivid_warp_render: colour = 8-bit LANCZOS down-sampling to
 * image_size, depth = centre sample -> project_depth(project_near, project_far) -> 256-entry uint8 RGB colour table `lut_host`
 * (cv2.COLORMAP_INFERNO pushed through colorize_depth's numpy steps).  Outputs uint8 [batch, H, W, 3] on the host. */
int ivid_warp_resolve_frame(ivid_warp_t* w, double project_near, double project_far, const uint8_t* lut_host, uint8_t* color8_host,
                            uint8_t* depth8_host, void* stream);

/* Training-pair warp — replaces rgbd_3d.SimpleRenderer (moderngl_renderer.py:11-148, shaders/simple.{vsh,fsh}) and
 * rgbd_3d.utils.forward_backward_warp (utils.py:335-417; called per training item by datasets/base.py:215-266).
 * SimpleRenderer.render(mesh, color, modelview, fov) for a single-sample handle: the mesh is an image_size^2 (padding=None)
 * or (image_size+2)^2 grid mesh in the 9-float vertex layout of ivid_warp_set_mesh (normals unused); outputs at render size
 * S on the host: color [S,S,3], depth [S,S] (linearised with the handle's near / far), mask [S,S] (alpha > 0.5 as 0/1). */
int ivid_warp_render_simple(ivid_warp_t* w, const float* verts_host, int nverts, const uint32_t* faces_host, int nfaces,
                            const float* color_host, const float* target_mv_host, double fov_deg, float* color_out_host,
                            float* depth_out_host, float* mask_out_host, void* stream);
/* forward_backward_warp for every sample of the handle's batch, device resident between the two renders (uses view slots 0 and
 * 1 of the handle and forgets any source views it held, like ivid_warp_reset):
 *   lin_depth0_host [batch,H,W] = linearize_depth(rgbd[..., 3:], near, far), color0_host [batch,H,W,3] = rgbd[..., :3];
 *   mv1 / mv0 [batch][16] (or one shared matrix each); params: padding (of the first mesh), fov, near, far, atol, rtol;
 *   out_host [batch,7,H,W]: color(3), depth, mask, mask, projected depth before masking. */
int ivid_warp_forward_backward(ivid_warp_t* w, const float* lin_depth0_host, const float* color0_host, const float* mv1_host,
                               const float* mv0_host, int shared_modelview, const ivid_warp_params_t* params, float* out_host,
                               void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Operator-level entry points (unit parity tests, profiling): the kernels the UNet is assembled from.
 * ------------------------------------------------------------------------------------------------------------------ */

/* nn.Conv2d 3x3 pad 1 / 1x1 as tcgen05 implicit GEMM.  act_dev fp16 NHWC [N,H,W,Cin] (Cin % 64 == 0); w_host fp32
 * [Cout,Cin,k,k] reference layout; optional 1x1 skip over act2_dev [N,H,W,Cin2] with w2_host [Cout,Cin2,1,1];
 * optional fp32 NHWC residual; out fp32 NHWC [N,H,W,Cout] (out_fp16 = 1: fp16). */
int ivid_op_conv2d(const void* act_dev, int N, int H, int W, int Cin, const float* w_host, const float* bias_host,
                   int Cout, int ksize, const void* act2_dev, int Cin2, const float* w2_host, const float* bias2_host,
                   const float* residual_dev, void* out_dev, int out_fp16, void* stream);
/* GroupNorm32 (+FiLM) (+SiLU) (+2x up / 2x2 avg-pool) over a virtual concat of two fp32 NHWC tensors -> fp16 NHWC. */
int ivid_op_group_norm(const float* x0_dev, int C0, const float* x1_dev, int C1, int N, int H, int W, int groups,
                       float eps, const float* gamma_host, const float* beta_host, const float* film_dev /*[N,2C]*/,
                       int silu, int mode, void* out_fp16_dev, void* stream);
/* QKVAttention (adm.py:233-253): qkv fp16 [N,T,3C] (legacy head-major q|k|v order) -> fp16 [N,T,C]. */
int ivid_op_attention(const void* qkv_dev, int N, int T, int C, void* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVID_B200_H_ */
