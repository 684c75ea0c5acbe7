// ADM UNet: topology / state-dict schema (host), packed device weights, per-batch execution plan.
// Mirrors the constructor logic of the reference's AdmUnet2d (diffusion/backbones/adm.py:318-487) so that the
// state-dict keys and shapes are identical, but executes the forward (adm.py:526-566) as a static list
// of sm_100a kernel launches over NHWC tensors.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ivid_b200.h"
#include "ops.h"

namespace ivid {

struct UnetConfig {
  int image_size = 0, in_channels = 0, model_channels = 0, out_channels = 0, num_res_blocks = 0;
  std::vector<int> attention_resolutions;
  std::vector<double> channel_mult{1, 2, 4, 8};
  bool conv_resample = true;
  int num_classes = 0;            // 0 = not class conditional
  bool has_null_class = false;
  bool use_fp16 = false;
  int num_groups = 32;
  int num_heads = 1;
  int num_head_channels = -1;
  bool use_scale_shift_norm = true;
  bool resblock_updown = true;
  double dropout = 0.0;
};

struct ParamSpec {
  std::string name;
  std::vector<int64_t> shape;
  bool is_buffer = false;
  bool set = false;
  std::vector<float> host;
  size_t numel() const { size_t n = 1; for (auto d : shape) n *= static_cast<size_t>(d); return n; }
};

struct ConvW { int cout = 0, cout_pad = 0, K = 0; size_t w_off = 0, b_off = 0; };
struct GnW { int C = 0; size_t g_off = 0, b_off = 0; };
struct LinW { int O = 0, K = 0; size_t w_off = 0, b_off = 0; };

struct ResBlockDef {
  std::string pfx;
  int cin = 0, cout = 0;
  int mode = 0;            // 0 same, 1 up, 2 down
  bool skip_conv = false;
  int film_off = 0;        // column offset of this block's (scale|shift) in the FiLM table
  GnW gn1, gn2;
  ConvW conv1, conv2;      // conv2 holds out_layers.3 (+ skip_connection as extra K columns)
};
struct AttnBlockDef {
  std::string pfx;
  int C = 0;
  GnW gn;
  ConvW qkv, proj;
};
struct ResampleDef {       // plain Downsample2d / Upsample2d layer (resblock_updown=False, adm.py:60-117)
  std::string pfx;         // layer name ("input_blocks.3.0", "output_blocks.2.1")
  int C = 0;
  int mode = 0;            // 1 up (nearest 2x [+ conv 3x3]), 2 down (conv 3x3 stride 2, or AvgPool2d(2))
  bool conv = false;       // conv_resample
  ConvW w;                 // op / conv weights, packed as an ordinary 3x3 conv
};
struct LayerRef { int kind; int idx; };   // kind 1 = ResBlock, 2 = AttentionBlock, 3 = plain resampling layer
struct BlockDef { std::vector<LayerRef> layers; bool is_input = false; bool is_output = false; };

struct Plan;

// Optional replacement of the output head's last kernel (eps_gather_kernel): the sampler hands the forward a launcher that
// consumes the tap columns Y directly (head_step_kernel: eps of both guidance halves -> mix -> x_{t-1}), so that eps never
// goes to HBM and the update is the last node of the forward's CUDA graph.  `key` must change whenever anything the launcher
// bakes in (pointers, scalars) changes: it is part of the graph-cache key.
struct HeadHook {
  std::function<void(const float* Y, const float* bias, int N, int H, int W, int Co, int ldy, cudaStream_t s)> launch;
  uint64_t key = 0;
};

class Unet {
 public:
  explicit Unet(const std::string& cfg_json);
  ~Unet();

  const UnetConfig& cfg() const { return cfg_; }
  const std::vector<ParamSpec>& params() const { return params_; }
  void set_param(const std::string& name, const float* data, const int64_t* shape, int ndim);
  void finalize(int device);
  bool finalized() const { return arena_ != nullptr; }
  void* arena() const { return arena_; }
  size_t arena_bytes() const { return arena_bytes_; }
  int device() const { return device_; }

  // forward over a batch of N samples; x rows are read modulo Nx (CFG halves share x)
  void forward(const float* x, int Nx, const ivid_cond_t* cond, const int64_t* t, const int64_t* classes, float* eps,
               int N, cudaStream_t stream, const HeadHook* hook = nullptr);
  // whether the output head runs as the tap-column GEMM whose last kernel a HeadHook can replace
  bool can_fuse_head() const;
  // Device-resident Philox stream id (step counter) of the conditional-input noise: the sampler points this at its step
  // state so that consecutive denoising steps replay the same CUDA graph (a by-value stream id would change the key).
  void set_cond_stream_dev(const int* p) { cond_stream_dev_ = p; }
  void debug_tap(int N, const std::string& name, float* host_out, size_t capacity, int* C, int* H, int* W);
  // per-kernel-family timing of the forwards issued between begin/end (CUDA events around every launch)
  void profile_begin();
  std::string profile_end();    // JSON: {"label": {"launches", "ms", "flops", "bytes"}, ...}

 private:
  void build_topology();
  int add_param(const std::string& name, std::vector<int64_t> shape, bool is_buffer = false);
  const ParamSpec& P(const std::string& name) const;
  Plan* get_plan(int N, int slot);
  Plan* build_plan(int N);

  UnetConfig cfg_;
  int embed_dim_ = 0;
  std::vector<ParamSpec> params_;
  std::map<std::string, int> pindex_;
  std::vector<ResBlockDef> res_;
  std::vector<AttnBlockDef> attn_;
  std::vector<ResampleDef> resample_;
  std::vector<BlockDef> blocks_;     // input blocks (block 0 = input conv, no layers), middle, output blocks in order
  int film_total_ = 0;
  int in_ch_stem_ = 0;               // channels after the input conv
  int final_ch_ = 0;

  // packed weights
  ConvW in_conv_, out_conv_;
  ConvW out1x1_;                     // split-precision 1x1 form of the output conv (9*Co tap columns), see unet.cu
  bool out_split_ = false;
  GnW out_gn_;
  LinW te1_, te2_, film_;
  size_t freqs_off_ = 0, label_off_ = 0;
  uint8_t* arena_ = nullptr;
  size_t arena_bytes_ = 0;
  int device_ = -1;

  const int* cond_stream_dev_ = nullptr;
  cudaStream_t cap_stream_ = nullptr;
  cudaStream_t side_stream_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  struct ProfAgg { int launches = 0; double ms = 0, flops = 0, bytes = 0; };
  bool profile_ = false;
  std::map<std::string, ProfAgg> profile_acc_;
  std::string profile_ops_;
  std::vector<std::unique_ptr<Plan>> plans_;
  friend struct Plan;
};

}  // namespace ivid
