#include "unet.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "json_min.h"

namespace ivid {

// ==================================================================================================
// config
// ==================================================================================================
static UnetConfig parse_config(const std::string& text) {
  JsonValue j;
  try {
    j = JsonParser(text).parse();
  } catch (const std::exception& e) {
    throw Error(kErrInvalidArgument, e.what());
  }
  IVID_REQUIRE(j.kind == JsonValue::kObject, "backbone config must be a JSON object");
  // accept a whole reference config file ({"backbone": {"name", "args"}, ...}), its "backbone" object, or bare args
  if (j.has("backbone") && j.at("backbone").kind == JsonValue::kObject) j = JsonValue(j.at("backbone"));
  if (j.has("args") && j.at("args").kind == JsonValue::kObject) j = JsonValue(j.at("args"));
  UnetConfig c;
  auto geti = [&](const char* k, int& dst, bool required) {
    if (j.has(k)) dst = static_cast<int>(j.at(k).num);
    else IVID_REQUIRE(!required, std::string("backbone config: missing '") + k + "'");
  };
  auto getb = [&](const char* k, bool& dst) { if (j.has(k)) dst = j.at(k).kind == JsonValue::kBool ? j.at(k).b : j.at(k).num != 0; };
  geti("image_size", c.image_size, true);
  geti("in_channels", c.in_channels, true);
  geti("model_channels", c.model_channels, true);
  geti("out_channels", c.out_channels, true);
  geti("num_res_blocks", c.num_res_blocks, true);
  IVID_REQUIRE(j.has("attention_resolutions"), "backbone config: missing 'attention_resolutions'");
  for (auto& v : j.at("attention_resolutions").arr) c.attention_resolutions.push_back(static_cast<int>(v.num));
  if (j.has("channel_mult")) {
    c.channel_mult.clear();
    for (auto& v : j.at("channel_mult").arr) c.channel_mult.push_back(v.num);
  }
  if (j.has("dropout")) c.dropout = j.at("dropout").num;
  getb("conv_resample", c.conv_resample);
  geti("num_classes", c.num_classes, false);
  getb("has_null_class", c.has_null_class);
  if (c.num_classes == 0) c.has_null_class = false;       // adm.py:350
  getb("use_fp16", c.use_fp16);
  geti("num_groups", c.num_groups, false);
  geti("num_heads", c.num_heads, false);                  // null -> keeps default (adm.py:330); unused with head channels
  geti("num_head_channels", c.num_head_channels, false);
  getb("use_scale_shift_norm", c.use_scale_shift_norm);
  getb("resblock_updown", c.resblock_updown);
  return c;
}

Unet::Unet(const std::string& cfg_json) : cfg_(parse_config(cfg_json)) { build_topology(); }

int Unet::add_param(const std::string& name, std::vector<int64_t> shape, bool is_buffer) {
  ParamSpec p;
  p.name = name;
  p.shape = std::move(shape);
  p.is_buffer = is_buffer;
  pindex_[name] = static_cast<int>(params_.size());
  params_.push_back(std::move(p));
  return static_cast<int>(params_.size()) - 1;
}

const ParamSpec& Unet::P(const std::string& name) const {
  auto it = pindex_.find(name);
  if (it == pindex_.end()) throw Error(kErrState, "internal: unknown parameter " + name);
  const ParamSpec& p = params_[it->second];
  if (!p.set) throw Error(kErrState, "parameter '" + name + "' was never set (load_state_dict incomplete)");
  return p;
}

// State-dict schema + block structure.  Follows the constructor of the reference (adm.py:356-487) layer for layer so that
// key names ("input_blocks.3.0.in_layers.0.weight", ...) and shapes match torch's registration order.
void Unet::build_topology() {
  const UnetConfig& c = cfg_;
  IVID_REQUIRE(c.model_channels % 64 == 0, "model_channels must be a multiple of 64 (tensor-core K slab)");
  IVID_REQUIRE(c.in_channels <= 16, "in_channels must be <= 16");
  IVID_REQUIRE(c.num_groups >= 1 && c.num_groups <= 64, "num_groups must be in [1,64]");
  const int mc = c.model_channels;
  const int E = mc * 4;
  embed_dim_ = E;

  add_param("time_embed.0.freqs", {mc / 2}, /*is_buffer=*/true);
  add_param("time_embed.1.weight", {E, mc});
  add_param("time_embed.1.bias", {E});
  add_param("time_embed.3.weight", {E, E});
  add_param("time_embed.3.bias", {E});
  if (c.num_classes > 0) add_param("label_emb.weight", {c.num_classes, E});

  auto heads_ok = [&](int ch) {
    const int hc = c.num_head_channels == -1 ? ch / std::max(1, c.num_heads) : c.num_head_channels;
    if (hc != 64 || ch % 64 != 0)
      throw Error(kErrNotImplemented, "attention head width must be 64 channels (num_head_channels=64)");
  };
  auto add_res = [&](const std::string& pfx, int cin, int cout, int mode) {
    ResBlockDef r;
    r.pfx = pfx; r.cin = cin; r.cout = cout; r.mode = mode;
    IVID_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "channel counts must be multiples of 64");
    IVID_REQUIRE(cin % c.num_groups == 0 && cout % c.num_groups == 0, "num_groups must divide channels");
    add_param(pfx + ".in_layers.0.weight", {cin});
    add_param(pfx + ".in_layers.0.bias", {cin});
    add_param(pfx + ".in_layers.2.weight", {cout, cin, 3, 3});
    add_param(pfx + ".in_layers.2.bias", {cout});
    const int ew = c.use_scale_shift_norm ? 2 * cout : cout;          // adm.py:176
    add_param(pfx + ".emb_layers.1.weight", {ew, E});
    add_param(pfx + ".emb_layers.1.bias", {ew});
    add_param(pfx + ".out_layers.0.weight", {cout});
    add_param(pfx + ".out_layers.0.bias", {cout});
    add_param(pfx + ".out_layers.3.weight", {cout, cout, 3, 3});
    add_param(pfx + ".out_layers.3.bias", {cout});
    r.skip_conv = (cin != cout);
    if (r.skip_conv) {
      add_param(pfx + ".skip_connection.weight", {cout, cin, 1, 1});
      add_param(pfx + ".skip_connection.bias", {cout});
    }
    r.film_off = film_total_;
    film_total_ += ew;
    res_.push_back(r);
    return static_cast<int>(res_.size()) - 1;
  };
  auto add_attn = [&](const std::string& pfx, int ch) {
    heads_ok(ch);
    AttnBlockDef a;
    a.pfx = pfx; a.C = ch;
    add_param(pfx + ".norm.weight", {ch});
    add_param(pfx + ".norm.bias", {ch});
    add_param(pfx + ".qkv.weight", {3 * ch, ch, 1});
    add_param(pfx + ".qkv.bias", {3 * ch});
    add_param(pfx + ".proj_out.weight", {ch, ch, 1});
    add_param(pfx + ".proj_out.bias", {ch});
    attn_.push_back(a);
    return static_cast<int>(attn_.size()) - 1;
  };
  auto add_resample = [&](const std::string& pfx, int ch, int mode) {
    ResampleDef r;
    r.pfx = pfx; r.C = ch; r.mode = mode; r.conv = c.conv_resample;
    IVID_REQUIRE(ch % 64 == 0, "channel counts must be multiples of 64");
    if (r.conv) {
      const std::string sub = mode == 2 ? ".op" : ".conv";      // Downsample2d.op (adm.py:111) / Upsample2d.conv (adm.py:81)
      add_param(pfx + sub + ".weight", {ch, ch, 3, 3});
      add_param(pfx + sub + ".bias", {ch});
    }
    resample_.push_back(r);
    return static_cast<int>(resample_.size()) - 1;
  };
  auto in_attn = [&](int ds) {
    return std::find(c.attention_resolutions.begin(), c.attention_resolutions.end(), ds) != c.attention_resolutions.end();
  };

  int ch = static_cast<int>(c.channel_mult[0] * mc);
  const int input_ch = ch;
  in_ch_stem_ = ch;
  add_param("input_blocks.0.0.weight", {ch, c.in_channels, 3, 3});
  add_param("input_blocks.0.0.bias", {ch});
  {
    BlockDef b;
    b.is_input = true;   // the stem conv: no layers, handled explicitly
    blocks_.push_back(b);
  }
  int ds = c.image_size;
  std::vector<int> input_block_chs{ch};
  int ib = 1;
  const int levels = static_cast<int>(c.channel_mult.size());
  for (int level = 0; level < levels; ++level) {
    const int outc = static_cast<int>(c.channel_mult[level] * mc);
    for (int r = 0; r < c.num_res_blocks; ++r) {
      BlockDef b;
      b.is_input = true;
      const std::string pfx = "input_blocks." + std::to_string(ib);
      b.layers.push_back({1, add_res(pfx + ".0", ch, outc, 0)});
      ch = outc;
      if (in_attn(ds)) b.layers.push_back({2, add_attn(pfx + ".1", ch)});
      blocks_.push_back(b);
      input_block_chs.push_back(ch);
      ++ib;
    }
    if (level != levels - 1) {
      BlockDef b;
      b.is_input = true;
      if (c.resblock_updown) b.layers.push_back({1, add_res("input_blocks." + std::to_string(ib) + ".0", ch, ch, 2)});
      else b.layers.push_back({3, add_resample("input_blocks." + std::to_string(ib) + ".0", ch, 2)});      // adm.py:409-412
      blocks_.push_back(b);
      input_block_chs.push_back(ch);
      ++ib;
      ds /= 2;
    }
  }
  {
    BlockDef b;
    b.layers.push_back({1, add_res("middle_block.0", ch, ch, 0)});
    b.layers.push_back({2, add_attn("middle_block.1", ch)});
    b.layers.push_back({1, add_res("middle_block.2", ch, ch, 0)});
    blocks_.push_back(b);
  }
  int ob = 0;
  for (int level = levels - 1; level >= 0; --level) {
    const int outc = static_cast<int>(mc * c.channel_mult[level]);
    for (int i = 0; i <= c.num_res_blocks; ++i) {
      BlockDef b;
      b.is_output = true;
      const std::string pfx = "output_blocks." + std::to_string(ob);
      const int ich = input_block_chs.back();
      input_block_chs.pop_back();
      int li = 0;
      b.layers.push_back({1, add_res(pfx + "." + std::to_string(li++), ch + ich, outc, 0)});
      ch = outc;
      if (in_attn(ds)) b.layers.push_back({2, add_attn(pfx + "." + std::to_string(li++), ch)});
      if (level != 0 && i == c.num_res_blocks) {
        if (c.resblock_updown) b.layers.push_back({1, add_res(pfx + "." + std::to_string(li++), ch, ch, 1)});
        else b.layers.push_back({3, add_resample(pfx + "." + std::to_string(li++), ch, 1)});               // adm.py:475-478
        ds *= 2;
      }
      blocks_.push_back(b);
      ++ob;
    }
  }
  final_ch_ = ch;
  IVID_REQUIRE(ch == input_ch, "final channel count must equal the stem width (adm.py:486 uses input_ch)");
  add_param("out.0.weight", {ch});
  add_param("out.0.bias", {ch});
  add_param("out.2.weight", {c.out_channels, input_ch, 3, 3});
  add_param("out.2.bias", {c.out_channels});
}

void Unet::set_param(const std::string& name, const float* data, const int64_t* shape, int ndim) {
  auto it = pindex_.find(name);
  if (it == pindex_.end()) throw Error(kErrInvalidArgument, "unexpected key in state_dict: " + name);
  ParamSpec& p = params_[it->second];
  bool same = static_cast<int>(p.shape.size()) == ndim;
  for (int i = 0; same && i < ndim; ++i) same = p.shape[i] == shape[i];
  if (!same) throw Error(kErrInvalidArgument, "size mismatch for " + name);
  p.host.assign(data, data + p.numel());
  p.set = true;
}

// ==================================================================================================
// weight packing
// ==================================================================================================
namespace {
struct ArenaBuilder {
  std::vector<uint8_t> buf;
  size_t alloc(size_t bytes) {
    const size_t off = (buf.size() + 255) & ~size_t(255);
    buf.resize(off + bytes, 0);
    return off;
  }
  template <class T> T* at(size_t off) { return reinterpret_cast<T*>(buf.data() + off); }
};

// [Cout][Cin][k][k] fp32 -> rows of [taps][cin_pad] fp16 written at column `kcol0` of a [cout_pad][Ktot] matrix
void pack_conv_rows(__half* dst, int Ktot, int kcol0, const float* w, int cout, int cin, int cin_pad, int ksz) {
  const int taps = ksz * ksz;
  for (int co = 0; co < cout; ++co)
    for (int tap = 0; tap < taps; ++tap)
      for (int ci = 0; ci < cin; ++ci)
        dst[static_cast<size_t>(co) * Ktot + kcol0 + tap * cin_pad + ci] =
            __float2half_rn(w[(static_cast<size_t>(co) * cin + ci) * taps + tap]);
}
// Stem: the 64 operand channels of the packed network input are  hi | lo | hi  of a two-term fp16 split of x
// (pack_input_kernel); the matching weight columns are  Wh | Wh | Wl  with W = Wh + Wl, so that the fp16 tensor-core
// product reproduces x*W to ~2^-21 (the lo*Wl term is dropped).
void pack_stem_rows(__half* dst, int Ktot, const float* w, int cout, int cin, int cin_pad) {
  for (int co = 0; co < cout; ++co)
    for (int tap = 0; tap < 9; ++tap)
      for (int ci = 0; ci < cin; ++ci) {
        const float v = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap];
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        __half* row = dst + static_cast<size_t>(co) * Ktot + tap * cin_pad;
        row[ci] = hi; row[cin + ci] = hi; row[2 * cin + ci] = lo;
      }
}

// Output head as a 1x1 GEMM over 9*Co columns (column tap*Co + c holds W[c][:, tap]), split precision:
// K = [Wh | Wh | Wl] against the activation segments [a_hi | a_lo | a_hi] (GnApplyParams::out_lo).
void pack_out_rows(__half* dst, int C, const float* w, int co_n) {
  const int K = 3 * C;
  for (int tap = 0; tap < 9; ++tap)
    for (int c = 0; c < co_n; ++c)
      for (int ci = 0; ci < C; ++ci) {
        const float v = w[(static_cast<size_t>(c) * C + ci) * 9 + tap];
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        __half* row = dst + static_cast<size_t>(tap * co_n + c) * K;
        row[ci] = hi; row[C + ci] = hi; row[2 * C + ci] = lo;
      }
}
}  // namespace

void Unet::finalize(int device) {
  IVID_CHECK_CUDA(cudaSetDevice(device));
  for (const auto& p : params_)
    if (!p.set) throw Error(kErrState, "parameter '" + p.name + "' was never set (load_state_dict incomplete)");
  plans_.clear();
  if (arena_) { cudaFree(arena_); arena_ = nullptr; }
  device_ = device;
  ArenaBuilder ab;
  auto put_f32 = [&](const std::vector<float>& v, size_t pad_to = 0) {
    const size_t n = std::max(v.size(), pad_to);
    const size_t off = ab.alloc(n * 4);
    std::memcpy(ab.at<float>(off), v.data(), v.size() * 4);
    return off;
  };
  auto pack_gn = [&](const std::string& pfx, int C) {
    GnW g; g.C = C;
    g.g_off = put_f32(P(pfx + ".weight").host);
    g.b_off = put_f32(P(pfx + ".bias").host);
    return g;
  };
  // generic conv: main weight (ksz x ksz over cin, channel-padded to cin_pad) + optional 1x1 skip weight over cin2
  auto pack_conv = [&](const std::string& wname, const std::string& bname, int cout, int cin, int cin_pad, int ksz,
                       const std::string& skip_w, const std::string& skip_b, int cin2) {
    ConvW cw;
    cw.cout = cout;
    cw.cout_pad = conv_pad_cout(cout);
    cw.K = ksz * ksz * cin_pad + cin2;
    cw.w_off = ab.alloc(static_cast<size_t>(cw.cout_pad) * cw.K * 2);
    pack_conv_rows(ab.at<__half>(cw.w_off), cw.K, 0, P(wname).host.data(), cout, cin, cin_pad, ksz);
    std::vector<float> bias(cw.cout_pad, 0.f);
    const auto& b = P(bname).host;
    for (int i = 0; i < cout; ++i) bias[i] = b[i];
    if (cin2 > 0) {
      pack_conv_rows(ab.at<__half>(cw.w_off), cw.K, ksz * ksz * cin_pad, P(skip_w).host.data(), cout, cin2, cin2, 1);
      const auto& b2 = P(skip_b).host;
      for (int i = 0; i < cout; ++i) bias[i] += b2[i];
    }
    cw.b_off = put_f32(bias);
    return cw;
  };
  auto pack_lin = [&](const std::string& pfx) {
    const ParamSpec& w = P(pfx + ".weight");
    LinW l; l.O = static_cast<int>(w.shape[0]); l.K = static_cast<int>(w.shape[1]);
    l.w_off = put_f32(w.host);
    l.b_off = put_f32(P(pfx + ".bias").host);
    return l;
  };

  freqs_off_ = put_f32(P("time_embed.0.freqs").host);
  te1_ = pack_lin("time_embed.1");
  te2_ = pack_lin("time_embed.3");
  if (cfg_.num_classes > 0) label_off_ = put_f32(P("label_emb.weight").host);
  // FiLM table weights: all emb_layers.1 stacked row-wise in ResBlock creation order
  {
    film_.O = film_total_; film_.K = embed_dim_;
    film_.w_off = ab.alloc(static_cast<size_t>(film_total_) * embed_dim_ * 4);
    film_.b_off = ab.alloc(static_cast<size_t>(film_total_) * 4);
    for (const auto& r : res_) {
      const auto& w = P(r.pfx + ".emb_layers.1.weight").host;
      const auto& b = P(r.pfx + ".emb_layers.1.bias").host;
      // swizzled K-chunk-major [E/32][film_total][32] when E % 32 == 0 (see film_table_kernel), row-major otherwise
      if (embed_dim_ % 32 == 0) {
        float* dst = ab.at<float>(film_.w_off);
        const int rows_w = static_cast<int>(w.size() / embed_dim_);
        for (int o = 0; o < rows_w; ++o)
          for (int kc = 0; kc < embed_dim_; kc += 32)
            for (int g = 0; g < 8; ++g)      // 16-byte groups XOR-swizzled with the output index (film_table_kernel)
              std::memcpy(dst + (static_cast<size_t>(kc / 32) * film_total_ + r.film_off + o) * 32 + ((g ^ ((r.film_off + o) & 7)) << 2),
                          w.data() + static_cast<size_t>(o) * embed_dim_ + kc + g * 4, 16);
      } else {
        std::memcpy(ab.at<float>(film_.w_off) + static_cast<size_t>(r.film_off) * embed_dim_, w.data(), w.size() * 4);
      }
      std::memcpy(ab.at<float>(film_.b_off) + r.film_off, b.data(), b.size() * 4);
    }
  }
  in_conv_ = pack_conv("input_blocks.0.0.weight", "input_blocks.0.0.bias", in_ch_stem_, cfg_.in_channels, 64, 3, "", "", 0);
  IVID_REQUIRE(3 * cfg_.in_channels <= 64, "in_channels must be <= 21 (two-term input split inside 64 operand channels)");
  pack_stem_rows(ab.at<__half>(in_conv_.w_off), in_conv_.K, P("input_blocks.0.0.weight").host.data(), in_ch_stem_, cfg_.in_channels, 64);
  for (auto& r : resample_)
    if (r.conv) {
      const std::string sub = r.mode == 2 ? ".op" : ".conv";
      r.w = pack_conv(r.pfx + sub + ".weight", r.pfx + sub + ".bias", r.C, r.C, r.C, 3, "", "", 0);
    }
  for (auto& r : res_) {
    r.gn1 = pack_gn(r.pfx + ".in_layers.0", r.cin);
    r.conv1 = pack_conv(r.pfx + ".in_layers.2.weight", r.pfx + ".in_layers.2.bias", r.cout, r.cin, r.cin, 3, "", "", 0);
    r.gn2 = pack_gn(r.pfx + ".out_layers.0", r.cout);
    r.conv2 = pack_conv(r.pfx + ".out_layers.3.weight", r.pfx + ".out_layers.3.bias", r.cout, r.cout, r.cout, 3,
                        r.pfx + ".skip_connection.weight", r.pfx + ".skip_connection.bias", r.skip_conv ? r.cin : 0);
  }
  for (auto& a : attn_) {
    a.gn = pack_gn(a.pfx + ".norm", a.C);
    a.qkv = pack_conv(a.pfx + ".qkv.weight", a.pfx + ".qkv.bias", 3 * a.C, a.C, a.C, 1, "", "", 0);
    a.proj = pack_conv(a.pfx + ".proj_out.weight", a.pfx + ".proj_out.bias", a.C, a.C, a.C, 1, "", "", 0);
  }
  out_gn_ = pack_gn("out.0", final_ch_);
  out_conv_ = pack_conv("out.2.weight", "out.2.bias", cfg_.out_channels, final_ch_, final_ch_, 3, "", "", 0);
  // split-precision 1x1 form of the same conv (see pack_out_rows); bias is added by eps_gather_kernel
  out_split_ = 9 * cfg_.out_channels <= 64 && getenv("IVID_NO_OUTSPLIT") == nullptr;
  if (out_split_) {
    out1x1_.cout = 64; out1x1_.cout_pad = 64; out1x1_.K = 3 * final_ch_;
    out1x1_.w_off = ab.alloc(static_cast<size_t>(64) * out1x1_.K * 2);
    pack_out_rows(ab.at<__half>(out1x1_.w_off), final_ch_, P("out.2.weight").host.data(), cfg_.out_channels);
    out1x1_.b_off = put_f32(std::vector<float>(64, 0.f));
  }

  arena_bytes_ = (ab.buf.size() + 255) & ~size_t(255);
  IVID_CHECK_CUDA(cudaMalloc(&arena_, arena_bytes_));
  IVID_CHECK_CUDA(cudaMemcpy(arena_, ab.buf.data(), ab.buf.size(), cudaMemcpyHostToDevice));
}

Unet::~Unet() {
  if (side_stream_) { cudaStreamDestroy(side_stream_); cudaEventDestroy(ev_fork_); cudaEventDestroy(ev_join_); }
  plans_.clear();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  if (arena_) cudaFree(arena_);
}

// ==================================================================================================
// execution plan
// ==================================================================================================
struct Plan {
  int N = 0;
  int slot = 0;            // 0 = main stream, 1 = side stream (two half-batches run concurrently)
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  double* stats_base = nullptr;
  size_t stats_bytes = 0;
  struct OpRec {
    std::function<void(cudaStream_t)> fn;
    const char* label;      // kernel family (profile aggregation key)
    double flops;           // algorithmic FLOPs of this launch (tensor-core ops)
    double bytes;           // algorithmic HBM bytes of this launch (memory-bound ops)
    std::string note;       // shape description (per-op profile dump)
  };
  struct OpList {
    std::vector<OpRec> v;
    const char* label = "other"; double flops = 0, bytes = 0; std::string note;    // metadata of the next push
    void tag(const char* l, double f, double b, std::string n = "") { label = l; flops = f; bytes = b; note = std::move(n); }
    void push_back(std::function<void(cudaStream_t)> fn) {
      v.push_back(OpRec{std::move(fn), label, flops, bytes, note});
      label = "other"; flops = 0; bytes = 0; note.clear();
    }
  } ops;
  std::vector<ConvLaunch*> convs;
  std::vector<AttnLaunch*> attns;
  // per-call inputs
  const float* x = nullptr; int Nx = 0; ivid_cond_t cond{}; const int64_t* t = nullptr; const int64_t* classes = nullptr;
  float* eps = nullptr;
  const HeadHook* hook = nullptr;
  const int* cond_stream_dev = nullptr;
  // CUDA graphs of the whole forward (memset + ~215 launches), one per distinct set of per-call pointers / by-value inputs:
  // the launch records bake them in, so a replay is valid exactly when the key matches.
  struct GraphKey {
    const void* x; int Nx; const void* t; const void* classes; void* eps;
    int kind; const void* y; const void* mask; const void* mask_rgb; const void* noise; uint64_t seed; uint32_t stream_id;
    const void* stream_dev;
    uint64_t hook_key;
    bool operator==(const GraphKey& o) const {
      return hook_key == o.hook_key && x == o.x && Nx == o.Nx && t == o.t && classes == o.classes && eps == o.eps && kind == o.kind && y == o.y && mask == o.mask &&
             mask_rgb == o.mask_rgb && noise == o.noise && seed == o.seed && stream_id == o.stream_id && stream_dev == o.stream_dev;
    }
  };
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec; uint64_t last_use; };
  // named block outputs (debug taps: per-layer parity tests read them back after a forward)
  struct Tap { std::string name; const float* d32; const __half* d16; int C, H, W; };
  std::vector<Tap> taps;
  std::vector<GraphEntry> graphs;
  uint64_t runs = 0;
  uint64_t graph_hits = 0, graph_captures = 0;      // a caller whose pointers change on every call gains nothing from capturing
  ~Plan() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    for (auto* c : convs) conv_launch_destroy(c);
    for (auto* a : attns) attn_launch_destroy(a);
    if (ws) cudaFree(ws);
  }
};

namespace {
struct Act {          // fp32 NHWC residual-stream tensor with per-(n,channel) statistics
  float* data = nullptr;
  __half* d16 = nullptr;    // fp16 copy written by the producing conv (operand of the next GroupNorm / 1x1 skip conv)
  double* stats = nullptr;
  int id = -1;              // position in creation order (fp32 liveness table of build_plan)
  int C = 0, H = 0, W = 0;
};
struct Bump {
  uint8_t* base; size_t off = 0;
  explicit Bump(uint8_t* b) : base(b) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};
}  // namespace

Plan* Unet::get_plan(int N, int slot) {
  for (auto& p : plans_) if (p->N == N && p->slot == slot) return p.get();
  if (plans_.size() >= 4) {
    IVID_CHECK_CUDA(cudaDeviceSynchronize());     // the evicted plan's workspace may still be in use by queued kernels
    plans_.erase(plans_.begin());
  }
  plans_.emplace_back(build_plan(N));
  plans_.back()->slot = slot;
  return plans_.back().get();
}

Plan* Unet::build_plan(int N) {
  std::unique_ptr<Plan> plan(new Plan());
  plan->N = N;
  Plan* pl = plan.get();
  const int S = cfg_.image_size;
  const int G = cfg_.num_groups;
  const float eps = 1e-5f;
  auto W8 = [&](size_t off) { return arena_ + off; };
  auto Wf = [&](size_t off) { return reinterpret_cast<const float*>(arena_ + off); };

  // ---- scratch maxima (walk the topology once for sizes) ----
  size_t max_act16 = 0, max_raw16 = 0, max_f32 = 0, max_c = 0, max_qkv = 0, max_col = 0;
  {
    int res = S;
    auto upd_res = [&](const ResBlockDef& r) {
      const int ro = r.mode == 1 ? res * 2 : (r.mode == 2 ? res / 2 : res);
      max_act16 = std::max({max_act16, static_cast<size_t>(N) * ro * ro * r.cin, static_cast<size_t>(N) * ro * ro * r.cout});
      max_raw16 = std::max(max_raw16, static_cast<size_t>(N) * res * res * r.cin);
      max_f32 = std::max({max_f32, static_cast<size_t>(N) * ro * ro * r.cin, static_cast<size_t>(N) * ro * ro * r.cout});
      max_c = std::max({max_c, static_cast<size_t>(r.cin), static_cast<size_t>(r.cout)});
      res = ro;
    };
    for (const auto& b : blocks_)
      for (const auto& l : b.layers) {
        if (l.kind == 1) upd_res(res_[l.idx]);
        else if (l.kind == 3) {
          const auto& r = resample_[l.idx];
          if (r.mode == 2) { max_col = std::max(max_col, static_cast<size_t>(N) * (res / 2) * (res / 2) * 9 * r.C); res /= 2; }
          else { max_act16 = std::max(max_act16, static_cast<size_t>(N) * (res * 2) * (res * 2) * r.C); res *= 2; }
        } else {
          const auto& a = attn_[l.idx];
          max_act16 = std::max(max_act16, static_cast<size_t>(N) * res * res * a.C);
          max_qkv = std::max(max_qkv, static_cast<size_t>(N) * res * res * 3 * a.C);
          max_c = std::max(max_c, static_cast<size_t>(a.C));
        }
      }
    max_act16 = std::max(max_act16, static_cast<size_t>(N) * S * S * std::max(64, final_ch_));
  }

  // The same allocation sequence is run three times: once to find out which block outputs are ever read in fp32 (residual
  // adds, resampling blocks; everything else consumes the fp16 copy), once to size the workspace, once to build the launches.
  std::vector<char> need32;
  bool collect = true;
  auto layout = [&](uint8_t* base, bool create) -> size_t {
    Bump bump(base);
    int next_id = 0;
    // statistics arena is placed first so that its base is known while creating ops
    // (sized generously: every tensor needs N*C*16 bytes; bound by total params walk below)
    size_t stats_cap = 0;
    {
      int res = S;
      stats_cap += static_cast<size_t>(N) * in_ch_stem_ * 16 + 1024;
      for (const auto& b : blocks_)
        for (const auto& l : b.layers) {
          if (l.kind == 1) {
            const auto& r = res_[l.idx];
            stats_cap += 2 * (static_cast<size_t>(N) * r.cout * 16 + 1024);
            res = r.mode == 1 ? res * 2 : (r.mode == 2 ? res / 2 : res);
          } else if (l.kind == 3) {
            stats_cap += static_cast<size_t>(N) * resample_[l.idx].C * 16 + 1024;
          } else {
            stats_cap += static_cast<size_t>(N) * attn_[l.idx].C * 16 + 1024;
          }
        }
    }
    uint8_t* stats_base = static_cast<uint8_t*>(bump.take(stats_cap));
    size_t soff = 0;
    auto take_stats = [&](int C) -> double* {
      const size_t o = soff;
      soff += (static_cast<size_t>(N) * C * 16 + 255) & ~size_t(255);
      return stats_base ? reinterpret_cast<double*>(stats_base + o) : nullptr;
    };
    auto stats_ptr = [&](const Act& a) { return a.stats; };
    auto new_act = [&](int C, int H, int Wd) {
      Act a; a.C = C; a.H = H; a.W = Wd;
      a.id = next_id++;
      if (collect) need32.resize(a.id + 1, 0);
      if (conv_can_out16(C)) a.d16 = static_cast<__half*>(bump.take(static_cast<size_t>(N) * H * Wd * C * 2));
      a.stats = take_stats(C);
      return a;
    };

    // scratch
    void* s_in = bump.take(static_cast<size_t>(N) * S * S * 64 * 2);
    void* s_a1 = bump.take(max_act16 * 2);
    void* s_a2 = bump.take(max_act16 * 2);
    void* s_xh = bump.take(max_raw16 * 2);
    float* s_xr = static_cast<float*>(bump.take(max_f32 * 4));
    float* s_h = static_cast<float*>(bump.take(max_f32 * 4));
    void* s_ab = bump.take(static_cast<size_t>(N) * max_c * 2 * 8);      // float2 per (n, c); 2x slack for concat
    void* s_qkv = bump.take(std::max<size_t>(max_qkv, 1) * 2);
    void* s_col = bump.take(std::max<size_t>(max_col, 1) * 2);           // im2col operand of the stride-2 Downsample2d conv
    float* s_pe = static_cast<float*>(bump.take(static_cast<size_t>(N) * cfg_.model_channels * 4));
    float* s_e1 = static_cast<float*>(bump.take(static_cast<size_t>(N) * embed_dim_ * 4));
    float* s_emb = static_cast<float*>(bump.take(static_cast<size_t>(N) * embed_dim_ * 4));
    float* s_film = static_cast<float*>(bump.take(static_cast<size_t>(N) * film_total_ * 4));
    float* s_xt = static_cast<float*>(bump.take(static_cast<size_t>((N + 31) / 32) * embed_dim_ * 32 * 4));

    // fp32 storage of a block output is allocated by its producer, and only when some consumer reads it (or the fp16-only
    // epilogue is not available: no fp16 copy, residual add, statistics not fusable)
    auto alloc32 = [&](Act& a) { a.data = static_cast<float*>(bump.take(static_cast<size_t>(N) * a.H * a.W * a.C * 4)); };
    auto only16 = [&](const Act& a, bool has_residual) {
      return !collect && !need32[a.id] && a.d16 != nullptr && !has_residual && conv_can_fuse_stats(a.H, a.W);
    };
    auto use32 = [&](const Act& a) -> const float* {
      if (collect) need32[a.id] = 1;
      else if (create) IVID_REQUIRE(a.data != nullptr, "internal: fp32 tensor was not materialised");
      return a.data;
    };
    Act pending_stats; bool has_pending_stats = false;
    auto add_conv = [&](ConvDesc d, const Act* stats_of = nullptr, double k_alg = 0.0, double n_alg = 0.0) {
      // GroupNorm statistics of the output are accumulated in the conv epilogue whenever the tile geometry allows
      const bool fused = stats_of != nullptr && conv_can_fuse_stats(d.H, d.W);
      if (fused) d.stats = stats_of->stats;
      if (stats_of != nullptr && !fused) pending_stats = *stats_of;
      has_pending_stats = stats_of != nullptr && !fused;
      if (!create) return;
      ConvLaunch* l = conv_launch_create(d);
      pl->convs.push_back(l);
      // ALGORITHMIC work: operand padding (stem: 9*Cin of 9*64 columns) and split-precision segments do not count
      const double K = k_alg > 0.0 ? k_alg : static_cast<double>(d.taps0) * d.C0 + (d.C1 > 0 ? static_cast<double>(d.taps1) * d.C1 : 0.0) +
                       (d.C2 > 0 ? static_cast<double>(d.taps2) * d.C2 : 0.0);
      const double M = static_cast<double>(d.N) * d.H * d.W;
      const double Nalg = n_alg > 0.0 ? n_alg : static_cast<double>(d.cout);
      static const char* names[] = {"conv_gemm<16>", "conv_gemm<64>", "conv_gemm<128>", "conv_gemm<256>"};
      const int bn = conv_launch_bn(l);
      pl->ops.tag(names[bn == 256 ? 3 : bn == 128 ? 2 : bn == 64 ? 1 : 0], 2.0 * M * K * Nalg,
                  M * (d.C0 + d.C1 + d.C2) * 2 + M * d.cout * ((d.out_mode == 1 ? 2 : 4) + (d.out16 ? 2 : 0) + (d.residual ? 4 : 0)) + K * d.cout_pad * 2,
                  std::to_string(d.H) + "x" + std::to_string(d.W) + " " + std::to_string(d.C0) + (d.C1 ? "+" + std::to_string(d.C1) : "") + (d.C2 ? "+" + std::to_string(d.C2) : "") +
                      "->" + std::to_string(d.cout) + " k" + std::to_string(d.taps0) + (d.residual ? " res" : "") + (d.stats ? " stats" : "") +
                      (d.out_mode == 1 ? " f16" : "") + (d.out16 ? " +f16" : ""));
      pl->ops.push_back([l](cudaStream_t s) { conv_launch_run(l, s); });
    };
    auto add_stats = [&](const Act& a) {
      if (!create || !has_pending_stats) return;      // already fused into the producing conv
      has_pending_stats = false;
      const float* x = a.data; double* st = stats_ptr(a);
      IVID_REQUIRE(!create || x != nullptr, "internal: statistics pass over a tensor without fp32 storage");
      const int HW = a.H * a.W, C = a.C;
      pl->ops.tag("gn_stats", 0, static_cast<double>(N) * HW * C * 4);
      pl->ops.push_back([=](cudaStream_t s) { launch_gn_stats(x, st, N, HW, C, s); });
    };
    // GroupNorm coefficients are computed in the prologue of the apply kernel: add_coeff only records its inputs
    GnApplyDesc pend;
    auto add_coeff = [&](const Act& a0, const Act* a1, const GnW& g, int film_off) {
      pend = GnApplyDesc();
      pend.stats0 = stats_ptr(a0);
      pend.stats1 = a1 ? stats_ptr(*a1) : nullptr;
      pend.groups = G; pend.eps = eps;
      pend.gamma = Wf(g.g_off); pend.beta = Wf(g.b_off);
      pend.film = film_off >= 0 ? s_film : nullptr;
      pend.film_ld = film_total_; pend.film_off = std::max(film_off, 0);
      pend.film_add = film_off >= 0 && !cfg_.use_scale_shift_norm;
    };
    auto add_apply = [&](GnApplyDesc d) {
      if (!create) return;
      d.stats0 = pend.stats0; d.stats1 = pend.stats1; d.groups = pend.groups; d.eps = pend.eps; d.gamma = pend.gamma;
      d.beta = pend.beta; d.film = pend.film; d.film_ld = pend.film_ld; d.film_off = pend.film_off; d.film_add = pend.film_add;
      {
        const int Ho = d.mode == 1 ? d.H * 2 : (d.mode == 2 ? d.H / 2 : d.H);
        const double in_el = static_cast<double>(d.N) * d.H * d.W * (d.C0 + d.C1);
        const double out_el = static_cast<double>(d.N) * Ho * Ho * (d.C0 + d.C1);
        pl->ops.tag("gn_apply", 0, in_el * (d.x0_half ? 2 : 4) + out_el * 2 + (d.out_raw16 ? out_el * 2 : 0) + (d.out_raw32 ? out_el * 4 : 0),
                    std::to_string(d.H) + "x" + std::to_string(d.W) + " C" + std::to_string(d.C0) + (d.C1 ? "+" + std::to_string(d.C1) : "") +
                        " m" + std::to_string(d.mode) + (d.x0_half ? " h16" : "") + (d.out_raw16 ? " raw16" : "") + (d.out_raw32 ? " raw32" : ""));
      }
      pl->ops.push_back([d](cudaStream_t s) { launch_gn_apply(d, s); });
    };

    // GroupNorm folded into the consuming conv (conv_fold_ok): only the per-(sample, channel) coefficients are computed here
    // (gn_coeff_kernel -> s_ab); the conv reads the RAW fp16 tensor and applies silu(A x + B) in its operand path
    auto add_fold_coeff = [&](const GnApplyDesc& g) {
      if (!create) return;
      GnApplyDesc d = g;
      d.stats0 = pend.stats0; d.stats1 = pend.stats1; d.groups = pend.groups; d.eps = pend.eps; d.gamma = pend.gamma;
      d.beta = pend.beta; d.film = pend.film; d.film_ld = pend.film_ld; d.film_off = pend.film_off; d.film_add = pend.film_add;
      pl->ops.tag("gn_coeff", 0, static_cast<double>(d.N) * (d.C0 + d.C1) * 24, "C" + std::to_string(d.C0 + d.C1));
      pl->ops.push_back([d, s_ab](cudaStream_t s) { launch_gn_coeff(d, s_ab, s); });
    };

    // ---- embeddings ----
    if (create) {
      const float* freqs = Wf(freqs_off_);
      const int half = cfg_.model_channels / 2, mc = cfg_.model_channels, E = embed_dim_;
      const float *w1 = Wf(te1_.w_off), *b1 = Wf(te1_.b_off), *w2 = Wf(te2_.w_off), *b2 = Wf(te2_.b_off);
      const float* lab = cfg_.num_classes > 0 ? Wf(label_off_) : nullptr;
      const float *wf = Wf(film_.w_off), *bf = Wf(film_.b_off);
      const int FT = film_total_;
      pl->ops.tag("embed", 2.0 * N * E * (mc + E), 4.0 * E * (mc + E), "posenc+time_embed");
      pl->ops.push_back([=](cudaStream_t s) {
        launch_posenc(pl->t, N, freqs, half, s_pe, N, s);
        launch_linear(s_pe, w1, b1, s_e1, N, mc, E, 0, nullptr, nullptr, 1, s);
        launch_linear(s_e1, w2, b2, s_emb, N, E, E, 1, pl->classes ? lab : nullptr, pl->classes, N, s);
      });
      pl->taps.push_back({"emb", s_emb, nullptr, E, 1, 1});                 // time (+ class) embedding [N, E] (adm.py:545-555)
      pl->taps.push_back({"film", s_film, nullptr, FT, 1, 1});              // all emb_layers outputs [N, sum 2*Cout] (adm.py:174-177, 214)
      pl->ops.tag("embed", 2.0 * N * E * FT, 4.0 * E * FT, "film table O=" + std::to_string(FT));
      pl->ops.push_back([=](cudaStream_t s) {
        if (E % 32 == 0) launch_film_table(s_emb, wf, bf, s_xt, s_film, N, E, FT, s);
        else launch_linear(s_emb, wf, bf, s_film, N, E, FT, 1, nullptr, nullptr, 1, s);
      });
    }

    // ---- stem ----
    if (create) {
      const int Cin = cfg_.in_channels, HW = S * S;
      pl->ops.tag("pack_input", 0, static_cast<double>(N) * HW * (Cin * 4 + 128));
      pl->ops.push_back([=](cudaStream_t s) {
        if (pl->cond.kind == 0) {
          launch_pack_input(pl->x, s_in, N, pl->Nx, Cin, HW, s);
        } else {
          CondPackDesc cp;
          cp.x = pl->x; cp.y = pl->cond.y_dev; cp.mask = pl->cond.mask_dev; cp.mask_rgb = pl->cond.mask_rgb_dev;
          cp.noise = pl->cond.noise_dev; cp.out = s_in; cp.N = N; cp.Nx = pl->Nx; cp.H = S; cp.W = S;
          cp.kind = pl->cond.kind; cp.seed = pl->cond.seed; cp.stream = pl->cond.stream_id; cp.stream_dev = pl->cond_stream_dev;
          launch_cond_pack(cp, s);
        }
      });
    }
    Act cur = new_act(in_ch_stem_, S, S);
    {
      ConvDesc d;
      d.act0 = s_in; d.C0 = 64; d.taps0 = 9;
      d.weight = W8(in_conv_.w_off); d.cout_pad = in_conv_.cout_pad; d.cout = in_conv_.cout; d.bias = Wf(in_conv_.b_off);
      if (only16(cur, false)) { d.out = cur.d16; d.out_mode = 1; }
      else { alloc32(cur); d.out = cur.data; d.out16 = cur.d16; d.out_mode = 0; }
      d.ldc = cur.C; d.N = N; d.H = S; d.W = S;
      add_conv(d, &cur, 9.0 * cfg_.in_channels);
      add_stats(cur);
      if (create) pl->taps.push_back({"input_blocks.0.0", cur.data, cur.d16, cur.C, cur.H, cur.W});
    }
    std::vector<Act> skips;
    skips.push_back(cur);

    auto run_res = [&](const ResBlockDef& r, const Act& x0, const Act* x1) -> Act {
      const int Cin = x0.C + (x1 ? x1->C : 0);
      IVID_REQUIRE(Cin == r.cin, "internal: ResBlock input width mismatch at " + r.pfx);
      const int H = x0.H, Wd = x0.W;
      const int Ho = r.mode == 1 ? H * 2 : (r.mode == 2 ? H / 2 : H);
      const int Wo = r.mode == 1 ? Wd * 2 : (r.mode == 2 ? Wd / 2 : Wd);
      const bool identity = !r.skip_conv;
      // nearest-2x upsample of the identity skip is applied by the conv epilogue itself when the tile geometry allows it
      const bool res_up = identity && r.mode == 1 && x1 == nullptr && conv_can_res_up(Wo, r.cout);
      const bool need_xr = identity && (r.mode != 0 || x1 != nullptr) && !res_up;   // resampled / concatenated identity skip
      // GN1 + SiLU (+ resample) -> a1 ; raw fp16 copy for the 1x1 skip conv ; raw fp32 for resampled identity skip
      add_coeff(x0, x1, r.gn1, -1);
      GnApplyDesc g1;
      // same-resolution blocks read the fp16 copies their producers wrote (half the GroupNorm read traffic, and the 1x1 skip
      // conv takes them directly as K segments: no raw copy pass)
      const bool use16 = r.mode == 0 && !need_xr && x0.d16 != nullptr && (x1 == nullptr || x1->d16 != nullptr);
      if (use16) { g1.x0 = x0.d16; g1.x1 = x1 ? x1->d16 : nullptr; g1.x0_half = true; }
      else { g1.x0 = use32(x0); g1.x1 = x1 ? use32(*x1) : nullptr; }
      g1.C0 = x0.C; g1.C1 = x1 ? x1->C : 0;
      g1.N = N; g1.H = H; g1.W = Wd; g1.mode = r.mode; g1.silu = 1;
      g1.out_act = s_a1; g1.out_raw16 = (r.skip_conv && !use16) ? s_xh : nullptr; g1.out_raw32 = need_xr ? s_xr : nullptr;
      IVID_REQUIRE(!(r.skip_conv && r.mode != 0), "internal: up/down ResBlocks keep the channel count");
      const bool fold1 = use16 && x1 == nullptr && conv_fold_ok(N, Ho, Wo, r.conv1.cout_pad, false);
      if (fold1) add_fold_coeff(g1); else add_apply(g1);
      // conv1 -> h (fp32) ; stats
      bool h_half = false;
      Act h; h.C = r.cout; h.H = Ho; h.W = Wo; h.data = s_h;
      h.stats = take_stats(r.cout);
      {
        ConvDesc d;
        d.act0 = s_a1; d.C0 = r.cin; d.taps0 = 9;
        if (fold1) { d.act0 = x0.d16; d.fold_ab = s_ab; d.fold_C = r.cin; d.fold_off0 = 0; }
        d.weight = W8(r.conv1.w_off); d.cout_pad = r.conv1.cout_pad; d.cout = r.cout; d.bias = Wf(r.conv1.b_off);
        // the hidden tensor only feeds GroupNorm 2: stored as fp16 (half the epilogue and GN traffic); its statistics are
        // taken from the rounded values in the conv epilogue.  Tiny feature maps keep the fp32 + stats-kernel path.
        h_half = conv_can_fuse_stats(Ho, Wo);
        d.out = h.data; d.ldc = r.cout; d.out_mode = h_half ? 1 : 0; d.N = N; d.H = Ho; d.W = Wo;
        add_conv(d, &h);
        add_stats(h);
      }
      // GN2 * (1+scale) + shift, SiLU -> a2
      add_coeff(h, nullptr, r.gn2, r.film_off);
      GnApplyDesc g2;
      g2.x0 = h.data; g2.x0_half = h_half; g2.C0 = r.cout; g2.N = N; g2.H = Ho; g2.W = Wo; g2.mode = 0; g2.silu = 1;
      g2.out_act = s_a2;
      const bool fold2 = h_half && conv_fold_ok(N, Ho, Wo, r.conv2.cout_pad, res_up);
      if (fold2) add_fold_coeff(g2); else add_apply(g2);
      // conv2 (+ 1x1 skip as extra K) + residual -> out
      Act out = new_act(r.cout, Ho, Wo);
      {
        ConvDesc d;
        d.act0 = s_a2; d.C0 = r.cout; d.taps0 = 9;
        if (fold2) { d.act0 = h.data; d.fold_ab = s_ab; d.fold_C = r.cout; d.fold_off0 = 0; }
        if (r.skip_conv && use16) {
          d.act1 = x0.d16; d.C1 = x0.C; d.taps1 = 1;
          if (x1 != nullptr) { d.act2 = x1->d16; d.C2 = x1->C; d.taps2 = 1; }
        } else if (r.skip_conv) { d.act1 = s_xh; d.C1 = r.cin; d.taps1 = 1; }
        d.weight = W8(r.conv2.w_off); d.cout_pad = r.conv2.cout_pad; d.cout = r.cout; d.bias = Wf(r.conv2.b_off);
        if (identity) { d.residual = need_xr ? s_xr : use32(x0); d.ldr = r.cout; d.residual_up = res_up; }
        if (only16(out, identity)) { d.out = out.d16; d.out_mode = 1; }
        else { alloc32(out); d.out = out.data; d.out16 = out.d16; d.out_mode = 0; }
        d.ldc = r.cout; d.N = N; d.H = Ho; d.W = Wo;
        add_conv(d, &out);
        add_stats(out);
      }
      if (create) pl->taps.push_back({r.pfx, out.data, out.d16, out.C, out.H, out.W});
      return out;
    };
    auto run_attn = [&](const AttnBlockDef& a, const Act& x) -> Act {
      IVID_REQUIRE(x.C == a.C, "internal: attention width mismatch at " + a.pfx);
      const int T = x.H * x.W;
      add_coeff(x, nullptr, a.gn, -1);
      GnApplyDesc g;
      if (x.d16 != nullptr) { g.x0 = x.d16; g.x0_half = true; } else g.x0 = use32(x);
      g.C0 = a.C; g.N = N; g.H = x.H; g.W = x.W; g.mode = 0; g.silu = 0; g.out_act = s_a1;
      add_apply(g);
      {
        ConvDesc d;
        d.act0 = s_a1; d.C0 = a.C; d.taps0 = 1;
        d.weight = W8(a.qkv.w_off); d.cout_pad = a.qkv.cout_pad; d.cout = 3 * a.C; d.bias = Wf(a.qkv.b_off);
        d.out = s_qkv; d.ldc = 3 * a.C; d.out_mode = 1; d.N = N; d.H = x.H; d.W = x.W;
        add_conv(d);
      }
      if (create) {
        AttnLaunch* l = attn_launch_create(s_qkv, N, T, a.C, s_a2);
        pl->attns.push_back(l);
        pl->ops.tag("attention", 4.0 * N * (a.C / 64) * static_cast<double>(T) * T * 64, static_cast<double>(N) * T * a.C * 8,
                    "T=" + std::to_string(T) + " heads=" + std::to_string(a.C / 64));
        pl->ops.push_back([l](cudaStream_t s) { attn_launch_run(l, s); });
      }
      Act out = new_act(a.C, x.H, x.W);
      {
        ConvDesc d;
        d.act0 = s_a2; d.C0 = a.C; d.taps0 = 1;
        d.weight = W8(a.proj.w_off); d.cout_pad = a.proj.cout_pad; d.cout = a.C; d.bias = Wf(a.proj.b_off);
        d.residual = use32(x); d.ldr = a.C;
        alloc32(out);
        d.out = out.data; d.out16 = out.d16; d.ldc = a.C; d.out_mode = 0; d.N = N; d.H = x.H; d.W = x.W;
        add_conv(d, &out);
        add_stats(out);
      }
      if (create) pl->taps.push_back({a.pfx, out.data, out.d16, out.C, out.H, out.W});
      return out;
    };

    // plain Downsample2d / Upsample2d (resblock_updown=False): the raw block output is the conv operand (no norm in between)
    auto run_resample = [&](const ResampleDef& r, const Act& x) -> Act {
      IVID_REQUIRE(x.C == r.C, "internal: resampling layer width mismatch at " + r.pfx);
      const int Ho = r.mode == 1 ? x.H * 2 : x.H / 2, Wo = r.mode == 1 ? x.W * 2 : x.W / 2;
      Act out = new_act(r.C, Ho, Wo);
      alloc32(out);
      if (r.conv) {
        IVID_REQUIRE(!create || x.d16 != nullptr, "internal: resampling conv needs the fp16 copy of its input");
        const void* x16 = x.d16;
        const int H = x.H, Wd = x.W, C = r.C;
        ConvDesc d;
        if (r.mode == 2) {
          if (create) {
            pl->ops.tag("resample", 0, static_cast<double>(N) * Ho * Wo * 9 * C * 4, "im2col s2 " + std::to_string(H) + "x" + std::to_string(Wd) + " C" + std::to_string(C));
            pl->ops.push_back([=](cudaStream_t s) { launch_im2col_s2(x16, s_col, N, H, Wd, C, s); });
          }
          d.act0 = s_col; d.C0 = 9 * C; d.taps0 = 1;
        } else {
          if (create) {
            pl->ops.tag("resample", 0, static_cast<double>(N) * Ho * Wo * C * 2.5, "nearest 2x " + std::to_string(H) + "x" + std::to_string(Wd) + " C" + std::to_string(C));
            pl->ops.push_back([=](cudaStream_t s) { launch_upsample2x_h16(x16, s_a1, N, H, Wd, C, s); });
          }
          d.act0 = s_a1; d.C0 = C; d.taps0 = 9;
        }
        d.weight = W8(r.w.w_off); d.cout_pad = r.w.cout_pad; d.cout = C; d.bias = Wf(r.w.b_off);
        d.out = out.data; d.out16 = out.d16; d.out_mode = 0; d.ldc = C; d.N = N; d.H = Ho; d.W = Wo;
        add_conv(d, &out, 9.0 * C);
        add_stats(out);
      } else {
        const float* src = use32(x);
        if (create) {
          float* dst = out.data; void* dst16 = out.d16;
          const int H = x.H, Wd = x.W, C = r.C, mode = r.mode;
          pl->ops.tag("resample", 0, static_cast<double>(N) * (H * Wd + Ho * Wo * 1.5) * C * 4, (mode == 1 ? "nearest 2x f32 " : "avgpool f32 ") + std::to_string(H) + "x" + std::to_string(Wd));
          pl->ops.push_back([=](cudaStream_t s) { launch_resample_f32(src, dst, dst16, N, H, Wd, C, mode, s); });
          double* st = stats_ptr(out);
          pl->ops.tag("gn_stats", 0, static_cast<double>(N) * Ho * Wo * C * 4);
          pl->ops.push_back([=](cudaStream_t s) { launch_gn_stats(dst, st, N, Ho * Wo, C, s); });
        }
      }
      if (create) pl->taps.push_back({r.pfx, out.data, out.d16, out.C, out.H, out.W});
      return out;
    };

    for (size_t bi = 1; bi < blocks_.size(); ++bi) {
      const BlockDef& b = blocks_[bi];
      bool first = true;
      for (const auto& l : b.layers) {
        if (l.kind == 1) {
          if (b.is_output && first) {
            Act sk = skips.back();
            skips.pop_back();
            cur = run_res(res_[l.idx], cur, &sk);
          } else {
            cur = run_res(res_[l.idx], cur, nullptr);
          }
        } else if (l.kind == 3) {
          cur = run_resample(resample_[l.idx], cur);
        } else {
          cur = run_attn(attn_[l.idx], cur);
        }
        first = false;
      }
      if (b.is_input) skips.push_back(cur);
    }
    IVID_REQUIRE(skips.empty(), "internal: skip stack not consumed");

    // ---- output head: GN + SiLU + conv3x3 -> eps (fp32 NCHW) ----
    add_coeff(cur, nullptr, out_gn_, -1);
    GnApplyDesc go;
    const bool split_head = out_split_ && cur.d16 != nullptr;
    if (cur.d16 != nullptr) { go.x0 = cur.d16; go.x0_half = true; } else go.x0 = use32(cur);
    go.C0 = cur.C; go.N = N; go.H = S; go.W = S; go.mode = 0; go.silu = 1; go.out_act = s_a1;
    if (split_head) go.out_lo = s_a2;
    add_apply(go);
    if (split_head) {
      // 1x1 GEMM over 9*Co tap columns on [a_hi | a_lo | a_hi] x [Wh | Wh | Wl], then shift-and-add + bias (eps_gather_kernel):
      // each activation element is read once instead of nine times, and the product carries ~21 mantissa bits
      ConvDesc d;
      d.act0 = s_a1; d.C0 = cur.C; d.taps0 = 1;
      d.act1 = s_a2; d.C1 = cur.C; d.taps1 = 1;
      d.act2 = s_a1; d.C2 = cur.C; d.taps2 = 1;
      d.weight = W8(out1x1_.w_off); d.cout_pad = 64; d.cout = 64; d.bias = Wf(out1x1_.b_off);
      d.out = s_h; d.ldc = 64; d.out_mode = 0; d.N = N; d.H = S; d.W = S;
      add_conv(d, nullptr, 9.0 * cur.C, static_cast<double>(cfg_.out_channels));
      if (create) {
        const float* Y = s_h; const float* ob = Wf(out_conv_.b_off);
        const int Co = cfg_.out_channels;
        pl->ops.tag("eps_gather", 0, static_cast<double>(N) * S * S * (9.0 * Co * 4 + Co * 4));
        pl->ops.push_back([=](cudaStream_t s) {
          if (pl->hook != nullptr) pl->hook->launch(Y, ob, N, S, S, Co, 64, s);
          else launch_eps_gather(Y, ob, pl->eps, N, S, S, Co, 64, s);
        });
      }
    } else if (create) {
      ConvDesc d;
      d.act0 = s_a1; d.C0 = cur.C; d.taps0 = 9;
      d.weight = W8(out_conv_.w_off); d.cout_pad = out_conv_.cout_pad; d.cout = cfg_.out_channels; d.bias = Wf(out_conv_.b_off);
      d.out = nullptr; d.ldc = 0; d.out_mode = 2; d.N = N; d.H = S; d.W = S;
      ConvLaunch* l = conv_launch_create(d);
      pl->convs.push_back(l);
      // eps pointer is a per-call input: patched through the plan at run time
      pl->ops.tag("conv_gemm<16>", 2.0 * N * S * S * 9.0 * cur.C * cfg_.out_channels,
                  static_cast<double>(N) * S * S * (cur.C * 2 + cfg_.out_channels * 4));
      pl->ops.push_back([l, pl](cudaStream_t s) { conv_launch_run_out(l, pl->eps, s); });
    }
    if (create) {
      pl->stats_base = reinterpret_cast<double*>(stats_base);
      pl->stats_bytes = soff;
      IVID_REQUIRE(soff <= stats_cap, "internal: statistics arena overflow");
    }
    return bump.off;
  };

  layout(nullptr, false);
  collect = false;
  const size_t total = layout(nullptr, false);
  IVID_CHECK_CUDA(cudaMalloc(&pl->ws, total + 4096));
  pl->ws_bytes = total;
  layout(pl->ws, true);
  return plan.release();
}

bool Unet::can_fuse_head() const {
  return out_split_ && conv_can_out16(final_ch_) && cfg_.out_channels == 4 && cfg_.image_size % 4 == 0;
}

void Unet::forward(const float* x, int Nx, const ivid_cond_t* cond, const int64_t* t, const int64_t* classes, float* eps,
                   int N, cudaStream_t stream, const HeadHook* hook) {
  if (!finalized()) throw Error(kErrState, "AdmUnet2d: forward before .cuda()/finalize");
  IVID_REQUIRE(N >= 1 && Nx >= 1 && N % Nx == 0, "forward: N must be a positive multiple of Nx");
  // reference: "this model is not class-conditioned" (adm.py:540)
  IVID_REQUIRE(classes == nullptr || cfg_.num_classes > 0, "this model is not class-conditioned");
  IVID_CHECK_CUDA(cudaSetDevice(device_));
  ivid_cond_t cnd = cond ? *cond : ivid_cond_t{};
  const int expect_in = cnd.kind == 1 ? (cnd.mask_rgb_dev ? 10 : 9) : (cnd.kind == 2 ? 8 : cfg_.in_channels);
  IVID_REQUIRE(expect_in == cfg_.in_channels, "forward: conditional inputs do not match the model's in_channels");
  IVID_REQUIRE(hook == nullptr || can_fuse_head(), "forward: a head hook needs the tap-column output head");
  IVID_REQUIRE(hook != nullptr || eps != nullptr, "forward: eps output missing");

  // Two half-batches on two streams: the HBM-bound GroupNorm passes of one half overlap the tensor-bound convolutions
  // of the other (a persistent conv CTA leaves enough registers / shared memory on every SM for a gn_apply block).
  // The halves are independent samples (typically the two classifier-free-guidance halves sharing x).
  // Measured on B200 (api.cu: ivid_debug_overlap): the two kernels do NOT overlap today (conv alone 0.21 ms + gn alone
  // 0.08 ms = 0.28 ms when issued concurrently), so the split is opt-in (IVID_SPLIT_BATCH=1) until the co-residency
  // blocker is understood.
  static const bool split_ok = getenv("IVID_SPLIT_BATCH") != nullptr;
  const bool can_split = split_ok && hook == nullptr && !profile_ && N % 2 == 0 && N >= 4 && (Nx == N || Nx == N / 2) &&
                         !(cnd.kind != 0 && Nx == N && cnd.noise_dev == nullptr);
  if (can_split) {
    const int half = N / 2;
    const size_t img = static_cast<size_t>(cfg_.image_size) * cfg_.image_size;
    if (side_stream_ == nullptr) {
      IVID_CHECK_CUDA(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
      IVID_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
      IVID_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    }
    IVID_CHECK_CUDA(cudaEventRecord(ev_fork_, stream));
    IVID_CHECK_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
    for (int hb = 0; hb < 2; ++hb) {
      Plan* ph = get_plan(half, hb);
      cudaStream_t st = hb == 0 ? stream : side_stream_;
      const bool shift = (Nx == N) && hb == 1;            // rows [half, N) of per-sample inputs
      const size_t xs = static_cast<size_t>(cfg_.in_channels) * img;
      ph->x = x + (shift && cnd.kind == 0 ? static_cast<size_t>(half) * xs : 0);
      if (cnd.kind != 0) ph->x = x + (shift ? static_cast<size_t>(half) * cfg_.out_channels * img : 0);
      ph->Nx = half;
      ph->t = t + hb * half;
      ph->classes = classes ? classes + hb * half : nullptr;
      ph->eps = eps + static_cast<size_t>(hb) * half * cfg_.out_channels * img;
      ph->cond = cnd;
      if (cnd.kind != 0 && shift) {
        const size_t yimg = cnd.kind == 2 ? img / 4 : img;
        ph->cond.y_dev = cnd.y_dev + static_cast<size_t>(half) * 4 * yimg;
        if (cnd.mask_dev) ph->cond.mask_dev = cnd.mask_dev + static_cast<size_t>(half) * img;
        if (cnd.mask_rgb_dev) ph->cond.mask_rgb_dev = cnd.mask_rgb_dev + static_cast<size_t>(half) * img;
        if (cnd.noise_dev) ph->cond.noise_dev = cnd.noise_dev + static_cast<size_t>(half) * 4 * img;
      }
      IVID_CHECK_CUDA(cudaMemsetAsync(ph->stats_base, 0, ph->stats_bytes, st));
      for (auto& op : ph->ops.v) op.fn(st);
    }
    IVID_CHECK_CUDA(cudaEventRecord(ev_join_, side_stream_));
    IVID_CHECK_CUDA(cudaStreamWaitEvent(stream, ev_join_, 0));
    return;
  }

  Plan* pl = get_plan(N, 0);
  pl->x = x; pl->Nx = Nx; pl->t = t; pl->classes = classes; pl->eps = eps;
  pl->cond = cnd;
  pl->cond_stream_dev = cond_stream_dev_;
  pl->hook = hook;
  if (!profile_) {
    // The first call of a plan runs eagerly (one-time function attributes, module loading); from the second call on the
    // forward is ONE cudaGraphLaunch.  Graphs are captured on a private stream (the caller's may be the legacy default
    // stream, which cannot be captured) and launched on the caller's stream.
    static const bool graphs_on = getenv("IVID_NO_GRAPH") == nullptr;
    ++pl->runs;
    if (graphs_on && pl->runs > 1) {
      const Plan::GraphKey key{x, Nx, t, classes, eps, cnd.kind, cnd.y_dev, cnd.mask_dev, cnd.mask_rgb_dev, cnd.noise_dev,
                               cnd.kind != 0 ? cnd.seed : 0ull, cnd.kind != 0 ? cnd.stream_id : 0u, cond_stream_dev_,
                               hook ? hook->key : 0ull};
      for (auto& g : pl->graphs)
        if (g.key == key) {
          g.last_use = pl->runs;
          ++pl->graph_hits;
          IVID_CHECK_CUDA(cudaGraphLaunch(g.exec, stream));
          return;
        }
      // capture pays off when keys repeat (the sampler loop: 2 keys); if 32 captures saw fewer hits than captures the caller
      // hands fresh buffers to every call (e.g. a Python loop that keeps every x_{t-1}): replay the launches on the stream
      const bool thrash = pl->graph_captures >= 32 && pl->graph_hits < pl->graph_captures;
      if (thrash) {
        IVID_CHECK_CUDA(cudaMemsetAsync(pl->stats_base, 0, pl->stats_bytes, stream));
        for (auto& op : pl->ops.v) op.fn(stream);
        return;
      }
      ++pl->graph_captures;
      if (cap_stream_ == nullptr) IVID_CHECK_CUDA(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
      cudaGraph_t graph = nullptr;
      IVID_CHECK_CUDA(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeRelaxed));
      try {
        IVID_CHECK_CUDA(cudaMemsetAsync(pl->stats_base, 0, pl->stats_bytes, cap_stream_));
        for (auto& op : pl->ops.v) op.fn(cap_stream_);
      } catch (...) {
        cudaStreamEndCapture(cap_stream_, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      IVID_CHECK_CUDA(cudaStreamEndCapture(cap_stream_, &graph));
      cudaGraphExec_t exec = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      IVID_CHECK_CUDA(ie);
      if (pl->graphs.size() >= 16) {
        // evict the least recently used graph; it may still be executing on the caller's stream
        size_t victim = 0;
        for (size_t i = 1; i < pl->graphs.size(); ++i) if (pl->graphs[i].last_use < pl->graphs[victim].last_use) victim = i;
        IVID_CHECK_CUDA(cudaStreamSynchronize(stream));
        cudaGraphExecDestroy(pl->graphs[victim].exec);
        pl->graphs.erase(pl->graphs.begin() + victim);
      }
      pl->graphs.push_back(Plan::GraphEntry{key, exec, pl->runs});
      IVID_CHECK_CUDA(cudaGraphLaunch(exec, stream));
      return;
    }
    IVID_CHECK_CUDA(cudaMemsetAsync(pl->stats_base, 0, pl->stats_bytes, stream));
    for (auto& op : pl->ops.v) op.fn(stream);
    return;
  }
  IVID_CHECK_CUDA(cudaMemsetAsync(pl->stats_base, 0, pl->stats_bytes, stream));
  // profiling pass: every launch bracketed by CUDA events on the launching stream (serialised; shares, not absolutes)
  std::vector<cudaEvent_t> ev(pl->ops.v.size() + 1);
  for (auto& e : ev) IVID_CHECK_CUDA(cudaEventCreate(&e));
  IVID_CHECK_CUDA(cudaEventRecord(ev[0], stream));
  for (size_t i = 0; i < pl->ops.v.size(); ++i) {
    pl->ops.v[i].fn(stream);
    IVID_CHECK_CUDA(cudaEventRecord(ev[i + 1], stream));
  }
  IVID_CHECK_CUDA(cudaStreamSynchronize(stream));
  for (size_t i = 0; i < pl->ops.v.size(); ++i) {
    float ms = 0.f;
    IVID_CHECK_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    auto& agg = profile_acc_[pl->ops.v[i].label];
    agg.launches += 1; agg.ms += ms; agg.flops += pl->ops.v[i].flops; agg.bytes += pl->ops.v[i].bytes;
    if (!pl->ops.v[i].note.empty()) {
      char buf[256];
      snprintf(buf, sizeof(buf), "[\"%s\", \"%s\", %.5f, %.4e, %.4e]", pl->ops.v[i].label, pl->ops.v[i].note.c_str(), ms, pl->ops.v[i].flops,
               pl->ops.v[i].bytes);
      if (!profile_ops_.empty()) profile_ops_ += ", ";
      profile_ops_ += buf;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
}

// Debug tap: output of the named layer (reference module path, e.g. "input_blocks.3.0") of the LAST forward of batch N,
// converted to fp32 NCHW on the host.  Tensors that only exist as fp16 in the plan (outputs nobody reads in fp32) are
// widened.  Synchronises the device; not on any hot path.
void Unet::debug_tap(int N, const std::string& name, float* host_out, size_t capacity, int* C, int* H, int* W) {
  Plan* pl = nullptr;
  for (auto& p : plans_) if (p->N == N && p->slot == 0) pl = p.get();
  if (pl == nullptr) throw Error(kErrState, "debug_tap: no forward of this batch size has run");
  for (const auto& t : pl->taps) {
    if (t.name != name) continue;
    const size_t n = static_cast<size_t>(N) * t.C * t.H * t.W;
    if (C) *C = t.C; if (H) *H = t.H; if (W) *W = t.W;
    if (host_out == nullptr) return;
    IVID_REQUIRE(capacity >= n, "debug_tap: output buffer too small");
    IVID_CHECK_CUDA(cudaSetDevice(device_));
    IVID_CHECK_CUDA(cudaDeviceSynchronize());
    std::vector<float> nhwc(n);
    if (t.d32 != nullptr) {
      IVID_CHECK_CUDA(cudaMemcpy(nhwc.data(), t.d32, n * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<__half> h(n);
      IVID_CHECK_CUDA(cudaMemcpy(h.data(), t.d16, n * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n; ++i) nhwc[i] = __half2float(h[i]);
    }
    const size_t HW = static_cast<size_t>(t.H) * t.W;
    for (int b = 0; b < N; ++b)
      for (size_t px = 0; px < HW; ++px)
        for (int c = 0; c < t.C; ++c) host_out[(static_cast<size_t>(b) * t.C + c) * HW + px] = nhwc[(static_cast<size_t>(b) * HW + px) * t.C + c];
    return;
  }
  throw Error(kErrInvalidArgument, "debug_tap: unknown layer '" + name + "'");
}

void Unet::profile_begin() { profile_ = true; profile_acc_.clear(); profile_ops_.clear(); }
std::string Unet::profile_end() {
  profile_ = false;
  std::string js = "{";
  bool first = true;
  for (const auto& kv : profile_acc_) {
    if (!first) js += ", ";
    first = false;
    char buf[256];
    snprintf(buf, sizeof(buf), "\"%s\": {\"launches\": %d, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", kv.first.c_str(),
             kv.second.launches, kv.second.ms, kv.second.flops, kv.second.bytes);
    js += buf;
  }
  if (getenv("IVID_PROFILE_OPS") != nullptr) js += ", \"_ops\": [" + profile_ops_ + "]";
  js += "}";
  return js;
}

}  // namespace ivid
