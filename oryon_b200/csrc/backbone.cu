// a1-a6 -- the network: Oryon.forward (reference net.py:142-167) for B (anchor, query) pairs.
//
//   encode_image   models/vlm.py:43-61     CLIP ViT-L/14@336 vision tower, ln_post patch tokens
//   encode_prompt  models/vlm.py:74-83     CLIP text tower on token ids (oryon_text_forward; cacheable per prompt set)
//   guidance       net.py:60-75            truncated torchvision swin_b -> 3 appearance-guidance maps
//   fusion         models/fusion.py:602-625
//   decoder        models/decoder.py:82-108
//
// Host side of the hot path: weight packing (fp16 split pairs, K-major, conv/convT weights permuted to the im2col
// column order, window maps, relative-position bias tables) and the op sequence.  Every matrix product runs on the
// tcgen05 GEMM (gemm.cu); everything else is in net_kernels.cu.  Activations are fp32 token-major (NHWC); the two
// images of every pair are processed together (N = 2B rows blocks: anchors first, then queries).
#include <cmath>
#include <cstdlib>
#include <map>
#include <string>

#include "gemm.cuh"
#include "net_kernels.cuh"

namespace oryon {
namespace attn {
int launch(oryon_handle* h, const __half* qkv_hi, const __half* qkv_lo, const __half* vt_hi, const __half* vt_lo, int ld_vt, int n_seq, int S,
           int heads, int width, int precision, __half* out_hi, __half* out_lo, int64_t ldh, cudaStream_t st, int out_lo_format = 0, int qk_f8x = 0);
bool v_from_qkv();
bool pp_active();
}
namespace net {

using gemm::round_up;

struct SplitW {  // GEMM-ready weight: [N][ld] fp16 split pair, ld = round_up(K, 64)
  __half* hi = nullptr;
  __half* lo = nullptr;
  int N = 0, K = 0, ld = 0;
  // precision-2 form (gemm.cuh): hi / lo hold W * 2^k (fp16 / 8-bit cross-term blocks), unscale = 2^-k goes into the epilogue's alpha
  int f8x = 0;
  float unscale = 1.f;
};

struct ClipBlock {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *qkv_b, *out_b, *fc_b, *proj_b;
  SplitW qkv, out, fc, proj;
};
struct SwinBlock {
  const float *n1_g, *n1_b, *n2_g, *n2_b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
  SplitW qkv, proj, fc1, fc2;
  const float* bias = nullptr;  // [heads][S][S] relative position bias (guidance backbone only)
  const float* bias_t = nullptr;  // the same table with the two token indices swapped (query index contiguous)
};
struct Stage {  // one resolution of a (guidance or fusion) Swin stack
  int H, W, ws, shift, heads, dim;
  int padH, padW, nW;
  const float* mask = nullptr;  // [nW][S][S] for the shifted block
};
struct UpW {
  SplitW up;                // ConvTranspose2d as a GEMM: [4*Cup][Cin]
  const float* up_b4;       // bias tiled over the 4 sub-pixels
  SplitW c0, c1;            // 3x3 convs (no bias)
  const float *g0_g, *g0_b, *g1_g, *g1_b;
  int cin, cup, cguid, cmid;
};

struct Backbone {
  oryon_backbone_config cfg{};
  bool finalized = false;
  std::map<std::string, std::vector<float>> host;   // until finalize
  std::vector<void*> allocs;

  // CLIP vision
  SplitW v_patch;
  const float *v_cls, *v_pos, *v_lnpre_g, *v_lnpre_b, *v_lnpost_g, *v_lnpost_b;
  std::vector<ClipBlock> v_blocks;
  // CLIP text
  const float *t_tok, *t_pos, *t_lnf_g, *t_lnf_b;
  std::vector<ClipBlock> t_blocks;
  SplitW t_proj;
  // guidance swin
  SplitW s_patch;
  const float *s_patch_b, *s_pn_g, *s_pn_b;
  SwinBlock s_blk[2][2];
  Stage s_stage[2];
  const float *s_mg_g[2], *s_mg_b[2];
  SplitW s_merge[2];
  // fusion
  SplitW f_clipconv, f_conv1, f_guid;
  const float *f_clipconv_b, *f_conv1_b, *f_guid_b, *f_text_w, *f_text_b;
  SwinBlock f_blk[2][2];
  const float *f_gn_g[2], *f_gn_b[2];
  Stage f_stage;
  ClassTfW f_ct[2];
  // decoder
  SplitW d_g0, d_g1;
  const float *d_g0_b, *d_g1_b;
  UpW d_up[3];
  const float *d_head_w, *d_head_b;

  // per-N window row maps (device), keyed by N
  std::map<int, std::vector<int32_t*>> maps;

  DeviceBuffer arena;
};

namespace {

constexpr int kVisW = 1024, kVisHeads = 16, kVisGrid = 24, kVisTok = 576, kVisPatchK = 588;
constexpr int kTxtW = 768, kTxtHeads = 12, kTxtL = 77, kVocab = 49408;
constexpr int kPrompts = 80;

struct Loader {
  oryon_handle* h;
  Backbone* m;
  cudaStream_t st;
  int rc = ORYON_OK;
  std::string missing;

  const std::vector<float>* get(const std::string& name, size_t numel) {
    auto it = m->host.find(name);
    if (it == m->host.end() || it->second.size() != numel) {
      if (rc == ORYON_OK) {
        rc = ORYON_ERR_NOT_LOADED;
        missing = name + (it == m->host.end() ? " (missing)" : " (wrong size)");
      }
      return nullptr;
    }
    return &it->second;
  }
  void* dmalloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) {
      if (rc == ORYON_OK) rc = ORYON_ERR_OUT_OF_MEMORY, missing = "cudaMalloc";
      return nullptr;
    }
    m->allocs.push_back(p);
    return p;
  }
  const float* f32(const std::vector<float>& v) {
    float* d = static_cast<float*>(dmalloc(v.size() * 4));
    if (d) cudaMemcpyAsync(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice, st), cudaStreamSynchronize(st);
    return d;
  }
  const float* f32(const std::string& name, size_t numel) {
    auto* v = get(name, numel);
    return v ? f32(*v) : nullptr;
  }
  // host matrix [N][K] (row-major) -> device split pair, K padded to 64.  f8x: the precision-2 form, for the layers that run with
  // fp8 cross terms (K % 64 == 0)
  SplitW split(const float* w, int N, int K, bool f8x = false) {
    SplitW s;
    s.N = N, s.K = K, s.ld = round_up(K, 64);
    float scale = 1.f;
    if (f8x && K % 64 == 0) {
      float amax = 0.f;
      for (size_t i = 0; i < (size_t)N * K; ++i) amax = std::max(amax, std::fabs(w[i]));
      scale = gemm::weight_scale(amax);
      s.f8x = 1, s.unscale = 1.f / scale;
    }
    float* tmp = nullptr;
    if (cudaMalloc(&tmp, (size_t)N * K * 4) != cudaSuccess) {
      if (rc == ORYON_OK) rc = ORYON_ERR_OUT_OF_MEMORY, missing = "cudaMalloc";
      return s;
    }
    s.hi = static_cast<__half*>(dmalloc((size_t)N * s.ld * 2));
    s.lo = static_cast<__half*>(dmalloc((size_t)N * s.ld * 2));
    if (s.hi && s.lo) {
      cudaMemcpyAsync(tmp, w, (size_t)N * K * 4, cudaMemcpyHostToDevice, st);
      const int r = s.f8x ? gemm::split_rows_f8x(h, tmp, K, N, K, s.hi, s.lo, s.ld, true, scale, st)
                          : gemm::split_rows(h, tmp, K, N, K, s.hi, s.lo, s.ld, st);
      if (r && rc == ORYON_OK) rc = r;
      cudaStreamSynchronize(st);
    }
    cudaFree(tmp);
    return s;
  }
  SplitW split(const std::string& name, int N, int K, bool f8x = false) {
    auto* v = get(name, (size_t)N * K);
    return v ? split(v->data(), N, K, f8x) : SplitW();
  }
  // conv weight [Cout][Cin][k][k] -> [Cout][(ky*k+kx)*Cin + ci]
  SplitW conv(const std::string& name, int cout, int cin, int k) {
    auto* v = get(name, (size_t)cout * cin * k * k);
    if (!v) return SplitW();
    std::vector<float> w((size_t)cout * cin * k * k);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < k * k; ++t) w[((size_t)co * k * k + t) * cin + ci] = (*v)[((size_t)co * cin + ci) * k * k + t];
    return split(w.data(), cout, cin * k * k);
  }
};

// window gather map: row (n, window, pos) of the padded + cyclically shifted grid -> source token row or -1 (padding)
std::vector<int32_t> window_map(int N, const Stage& s, int shift) {
  const int S = s.ws * s.ws, nwx = s.padW / s.ws;
  std::vector<int32_t> map((size_t)N * s.nW * S);
  for (int n = 0; n < N; ++n)
    for (int w = 0; w < s.nW; ++w)
      for (int p = 0; p < S; ++p) {
        const int py = (w / nwx) * s.ws + p / s.ws, px = (w % nwx) * s.ws + p % s.ws;
        const int sy = (py + shift) % s.padH, sx = (px + shift) % s.padW;
        map[((size_t)n * s.nW + w) * S + p] = (sy < s.H && sx < s.W) ? n * s.H * s.W + sy * s.W + sx : -1;
      }
  return map;
}

// shifted-window mask (0 / -100) on the padded, shifted grid (torchvision shifted_window_attention; fusion.py:147-167)
std::vector<float> shift_mask(const Stage& s) {
  const int S = s.ws * s.ws, nwx = s.padW / s.ws;
  std::vector<int> id((size_t)s.padH * s.padW);
  auto region = [&](int v, int extent) { return v < extent - s.ws ? 0 : (v < extent - s.shift ? 1 : 2); };
  for (int y = 0; y < s.padH; ++y)
    for (int x = 0; x < s.padW; ++x) id[(size_t)y * s.padW + x] = region(y, s.padH) * 3 + region(x, s.padW);
  std::vector<float> m((size_t)s.nW * S * S);
  for (int w = 0; w < s.nW; ++w)
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        const int yi = (w / nwx) * s.ws + i / s.ws, xi = (w % nwx) * s.ws + i % s.ws;
        const int yj = (w / nwx) * s.ws + j / s.ws, xj = (w % nwx) * s.ws + j % s.ws;
        m[((size_t)w * S + i) * S + j] = id[(size_t)yi * s.padW + xi] == id[(size_t)yj * s.padW + xj] ? 0.f : -100.f;
      }
  return m;
}

Stage make_stage(int H, int W, int ws, int heads, int dim) {
  Stage s;
  s.H = H, s.W = W, s.ws = ws, s.shift = ws / 2, s.heads = heads, s.dim = dim;
  s.padH = round_up(H, ws), s.padW = round_up(W, ws);
  s.nW = (s.padH / ws) * (s.padW / ws);
  return s;
}

struct Arena {
  char* base = nullptr;
  size_t off = 0, peak = 0;
  template <typename T>
  T* take(size_t count) {
    const size_t o = off;
    off += (count * sizeof(T) + 255) & ~size_t(255);
    peak = std::max(peak, off);
    return base ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

struct SplitA {  // activation operand
  __half* hi;
  __half* lo;
  int ld;
};

}  // namespace

static Backbone* model_of(oryon_handle* h) { return static_cast<Backbone*>(h->backbone); }

void destroy_backbone(oryon_handle* h) {
  Backbone* m = model_of(h);
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  for (auto& kv : m->maps)
    for (auto* p : kv.second) cudaFree(p);
  m->arena.release();
  delete m;
  h->backbone = nullptr;
}

int set_weight(oryon_handle* h, const char* name, const float* data, int64_t numel) {
  ORYON_REQUIRE(h && name && data && numel > 0, "oryon_backbone_set_weight: bad argument");
  if (!h->backbone) h->backbone = new Backbone();
  Backbone* m = model_of(h);
  ORYON_REQUIRE(!m->finalized, "oryon_backbone_set_weight: the model is finalized; destroy the handle to reload");
  m->host[name].assign(data, data + numel);
  return ORYON_OK;
}

static void load_clip_blocks(Loader& L, const std::string& prefix, int width, int layers, std::vector<ClipBlock>& out, bool f8x) {
  out.resize(layers);
  for (int i = 0; i < layers; ++i) {
    const std::string p = prefix + ".resblocks." + std::to_string(i);
    ClipBlock& b = out[i];
    b.ln1_g = L.f32(p + ".ln_1.weight", width), b.ln1_b = L.f32(p + ".ln_1.bias", width);
    b.ln2_g = L.f32(p + ".ln_2.weight", width), b.ln2_b = L.f32(p + ".ln_2.bias", width);
    b.qkv = L.split(p + ".attn.in_proj_weight", 3 * width, width, f8x), b.qkv_b = L.f32(p + ".attn.in_proj_bias", 3 * width);
    b.out = L.split(p + ".attn.out_proj.weight", width, width, f8x), b.out_b = L.f32(p + ".attn.out_proj.bias", width);
    b.fc = L.split(p + ".mlp.c_fc.weight", 4 * width, width, f8x), b.fc_b = L.f32(p + ".mlp.c_fc.bias", 4 * width);
    b.proj = L.split(p + ".mlp.c_proj.weight", width, 4 * width, f8x), b.proj_b = L.f32(p + ".mlp.c_proj.bias", width);
  }
}

static std::vector<float> transpose(const std::vector<float>& w, int rows, int cols) {
  std::vector<float> t(w.size());
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
  return t;
}

int finalize(oryon_handle* h, const oryon_backbone_config* cfg, cudaStream_t st) {
  ORYON_REQUIRE(h && cfg, "oryon_backbone_finalize: null argument");
  Backbone* m = model_of(h);
  if (!m) {
    set_error("oryon_backbone_finalize: no weights were set");
    return ORYON_ERR_NOT_LOADED;
  }
  ORYON_REQUIRE(!m->finalized, "oryon_backbone_finalize: already finalized");
  ORYON_REQUIRE(cfg->vis_layers >= 1 && cfg->txt_layers >= 0 && cfg->precision >= 1 && cfg->precision <= 3, "oryon_backbone_finalize: bad config");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  m->cfg = *cfg;
  Loader L{h, m, st};

  {  // CLIP vision (clip/model.py VisionTransformer as driven by vlm.py:46-56)
    const std::string p = "vlm.clip_model.visual";
    m->v_patch = L.split(p + ".conv1.weight", kVisW, kVisPatchK);  // [1024][3*14*14]: already the (c, ky, kx) flatten
    m->v_cls = L.f32(p + ".class_embedding", kVisW);
    m->v_pos = L.f32(p + ".positional_embedding", (size_t)(kVisTok + 1) * kVisW);
    m->v_lnpre_g = L.f32(p + ".ln_pre.weight", kVisW), m->v_lnpre_b = L.f32(p + ".ln_pre.bias", kVisW);
    m->v_lnpost_g = L.f32(p + ".ln_post.weight", kVisW), m->v_lnpost_b = L.f32(p + ".ln_post.bias", kVisW);
    // precision 2: the four linear layers of every vision block (96 of the 148 GEMMs, 9/10 of the GEMM time) run with fp8 cross terms
    load_clip_blocks(L, p + ".transformer", kVisW, cfg->vis_layers, m->v_blocks, cfg->precision == 2);
  }
  if (cfg->txt_layers > 0) {  // CLIP text
    const std::string p = "vlm.clip_model";
    m->t_tok = L.f32(p + ".token_embedding.weight", (size_t)kVocab * kTxtW);
    m->t_pos = L.f32(p + ".positional_embedding", (size_t)kTxtL * kTxtW);
    m->t_lnf_g = L.f32(p + ".ln_final.weight", kTxtW), m->t_lnf_b = L.f32(p + ".ln_final.bias", kTxtW);
    load_clip_blocks(L, p + ".transformer", kTxtW, cfg->txt_layers, m->t_blocks, false);
    if (auto* v = L.get(p + ".text_projection", (size_t)kTxtW * kTxtW)) {  // x @ P  ==  linear(x, P^T)
      const std::vector<float> t = transpose(*v, kTxtW, kTxtW);
      m->t_proj = L.split(t.data(), kTxtW, kTxtW);
    }
  }
  {  // guidance backbone (torchvision swin_b, features.0 .. features.4)
    const std::string p = "guidance_backbone.features";
    m->s_patch = L.split(p + ".0.0.weight", 128, 48);
    m->s_patch_b = L.f32(p + ".0.0.bias", 128);
    m->s_pn_g = L.f32(p + ".0.2.weight", 128), m->s_pn_b = L.f32(p + ".0.2.bias", 128);
    m->s_stage[0] = make_stage(96, 96, 7, 4, 128);
    m->s_stage[1] = make_stage(48, 48, 7, 8, 256);
    for (int si = 0; si < 2; ++si) {
      Stage& S = m->s_stage[si];
      const int dim = S.dim, heads = S.heads, feat = si == 0 ? 1 : 3;
      S.mask = L.f32(shift_mask(S));
      for (int bi = 0; bi < 2; ++bi) {
        const std::string b = p + "." + std::to_string(feat) + "." + std::to_string(bi);
        SwinBlock& B = m->s_blk[si][bi];
        B.n1_g = L.f32(b + ".norm1.weight", dim), B.n1_b = L.f32(b + ".norm1.bias", dim);
        B.n2_g = L.f32(b + ".norm2.weight", dim), B.n2_b = L.f32(b + ".norm2.bias", dim);
        B.qkv = L.split(b + ".attn.qkv.weight", 3 * dim, dim), B.qkv_b = L.f32(b + ".attn.qkv.bias", 3 * dim);
        B.proj = L.split(b + ".attn.proj.weight", dim, dim), B.proj_b = L.f32(b + ".attn.proj.bias", dim);
        B.fc1 = L.split(b + ".mlp.0.weight", 4 * dim, dim), B.fc1_b = L.f32(b + ".mlp.0.bias", 4 * dim);
        B.fc2 = L.split(b + ".mlp.3.weight", dim, 4 * dim), B.fc2_b = L.f32(b + ".mlp.3.bias", dim);
        if (auto* tab = L.get(b + ".attn.relative_position_bias_table", (size_t)169 * heads)) {
          // torchvision _get_relative_position_bias with the standard 7x7 index
          std::vector<float> bias((size_t)heads * 49 * 49);
          for (int i = 0; i < 49; ++i)
            for (int j = 0; j < 49; ++j) {
              const int dy = i / 7 - j / 7 + 6, dx = i % 7 - j % 7 + 6;
              for (int hd = 0; hd < heads; ++hd) bias[((size_t)hd * 49 + i) * 49 + j] = (*tab)[(size_t)(dy * 13 + dx) * heads + hd];
            }
          B.bias = L.f32(bias);
          std::vector<float> bias_t(bias.size());
          for (int hd = 0; hd < heads; ++hd)
            for (int i = 0; i < 49; ++i)
              for (int j = 0; j < 49; ++j) bias_t[((size_t)hd * 49 + j) * 49 + i] = bias[((size_t)hd * 49 + i) * 49 + j];
          B.bias_t = L.f32(bias_t);
        }
      }
      const std::string mg = p + "." + std::to_string(feat + 1);
      m->s_mg_g[si] = L.f32(mg + ".norm.weight", 4 * dim), m->s_mg_b[si] = L.f32(mg + ".norm.bias", 4 * dim);
      m->s_merge[si] = L.split(mg + ".reduction.weight", 2 * dim, 4 * dim);
    }
  }
  {  // fusion (models/fusion.py)
    const std::string p = "fusion";
    m->f_clipconv = L.split(p + ".clip_conv.weight", 768, 1024), m->f_clipconv_b = L.f32(p + ".clip_conv.bias", 768);
    m->f_conv1 = L.conv(p + ".conv1.weight", 128, kPrompts, 7), m->f_conv1_b = L.f32(p + ".conv1.bias", 128);
    m->f_guid = L.conv(p + ".guidance_projection.0.weight", 128, 512, 3), m->f_guid_b = L.f32(p + ".guidance_projection.0.bias", 128);
    m->f_text_w = L.f32(p + ".text_guidance_projection.0.weight", (size_t)128 * 768);
    m->f_text_b = L.f32(p + ".text_guidance_projection.0.bias", 128);
    m->f_stage = make_stage(24, 24, 12, 4, 128);
    m->f_stage.mask = L.f32(shift_mask(m->f_stage));
    for (int li = 0; li < 2; ++li) {
      const std::string lp = p + ".layers." + std::to_string(li);
      m->f_gn_g[li] = L.f32(lp + ".swin_block.guidance_norm.weight", 128), m->f_gn_b[li] = L.f32(lp + ".swin_block.guidance_norm.bias", 128);
      for (int bi = 0; bi < 2; ++bi) {
        const std::string b = lp + ".swin_block.block_" + std::to_string(bi + 1);
        SwinBlock& B = m->f_blk[li][bi];
        B.n1_g = L.f32(b + ".norm1.weight", 128), B.n1_b = L.f32(b + ".norm1.bias", 128);
        B.n2_g = L.f32(b + ".norm2.weight", 128), B.n2_b = L.f32(b + ".norm2.bias", 128);
        // q, k read cat(norm1(x), guidance) (256), v reads the first 128 columns only (fusion.py:83-85): one [384][256] GEMM
        auto *qw = L.get(b + ".attn.q.weight", 128 * 256), *kw = L.get(b + ".attn.k.weight", 128 * 256), *vw = L.get(b + ".attn.v.weight", 128 * 128);
        auto *qb = L.get(b + ".attn.q.bias", 128), *kb = L.get(b + ".attn.k.bias", 128), *vb = L.get(b + ".attn.v.bias", 128);
        if (qw && kw && vw && qb && kb && vb) {
          std::vector<float> w((size_t)384 * 256, 0.f), bb(384);
          std::copy(qw->begin(), qw->end(), w.begin());
          std::copy(kw->begin(), kw->end(), w.begin() + 128 * 256);
          for (int r = 0; r < 128; ++r) std::copy(vw->begin() + r * 128, vw->begin() + (r + 1) * 128, w.begin() + (size_t)(256 + r) * 256);
          std::copy(qb->begin(), qb->end(), bb.begin()), std::copy(kb->begin(), kb->end(), bb.begin() + 128),
              std::copy(vb->begin(), vb->end(), bb.begin() + 256);
          B.qkv = L.split(w.data(), 384, 256), B.qkv_b = L.f32(bb);
        }
        B.proj = L.split(b + ".attn.proj.weight", 128, 128), B.proj_b = L.f32(b + ".attn.proj.bias", 128);
        B.fc1 = L.split(b + ".mlp.fc1.weight", 512, 128), B.fc1_b = L.f32(b + ".mlp.fc1.bias", 512);
        B.fc2 = L.split(b + ".mlp.fc2.weight", 128, 512), B.fc2_b = L.f32(b + ".mlp.fc2.bias", 128);
      }
      const std::string a = lp + ".attention";
      ClassTfW& C = m->f_ct[li];
      C.n1_g = L.f32(a + ".norm1.weight", 128), C.n1_b = L.f32(a + ".norm1.bias", 128);
      C.n2_g = L.f32(a + ".norm2.weight", 128), C.n2_b = L.f32(a + ".norm2.bias", 128);
      auto tw = [&](const std::string& n, int rows, int cols) -> const float* {
        auto* v = L.get(n, (size_t)rows * cols);
        return v ? L.f32(transpose(*v, rows, cols)) : nullptr;
      };
      C.q_w = tw(a + ".attention.q.weight", 128, 256), C.q_b = L.f32(a + ".attention.q.bias", 128);
      C.k_w = tw(a + ".attention.k.weight", 128, 256), C.k_b = L.f32(a + ".attention.k.bias", 128);
      C.v_w = tw(a + ".attention.v.weight", 128, 128), C.v_b = L.f32(a + ".attention.v.bias", 128);
      C.m0_w = tw(a + ".MLP.0.weight", 512, 128), C.m0_b = L.f32(a + ".MLP.0.bias", 512);
      C.m2_w = tw(a + ".MLP.2.weight", 128, 512), C.m2_b = L.f32(a + ".MLP.2.bias", 128);
    }
  }
  {  // decoder (models/decoder.py)
    const std::string p = "decoder";
    m->d_g0 = L.conv(p + ".decoder_guidance_projection.0.0.weight", 32, 256, 3), m->d_g0_b = L.f32(p + ".decoder_guidance_projection.0.0.bias", 32);
    m->d_g1 = L.conv(p + ".decoder_guidance_projection.1.0.weight", 16, 128, 3), m->d_g1_b = L.f32(p + ".decoder_guidance_projection.1.0.bias", 16);
    const int spec[3][4] = {{128, 96, 32, 64}, {64, 48, 16, 32}, {32, 32, 0, 32}};  // cin, cup, cguid, cmid
    for (int i = 0; i < 3; ++i) {
      UpW& U = m->d_up[i];
      U.cin = spec[i][0], U.cup = spec[i][1], U.cguid = spec[i][2], U.cmid = spec[i][3];
      const std::string d = p + ".decoder" + std::to_string(i + 1);
      auto *uw = L.get(d + ".up.weight", (size_t)U.cin * U.cup * 4), *ub = L.get(d + ".up.bias", U.cup);
      if (uw && ub) {  // ConvTranspose2d weight [Cin][Cup][2][2] -> [(dy*2+dx)*Cup + co][ci]
        std::vector<float> w((size_t)4 * U.cup * U.cin), b4((size_t)4 * U.cup);
        for (int ci = 0; ci < U.cin; ++ci)
          for (int co = 0; co < U.cup; ++co)
            for (int t = 0; t < 4; ++t) w[((size_t)t * U.cup + co) * U.cin + ci] = (*uw)[((size_t)ci * U.cup + co) * 4 + t];
        for (int t = 0; t < 4; ++t) std::copy(ub->begin(), ub->end(), b4.begin() + (size_t)t * U.cup);
        U.up = L.split(w.data(), 4 * U.cup, U.cin), U.up_b4 = L.f32(b4);
      }
      U.c0 = L.conv(d + ".conv.double_conv.0.weight", U.cmid, U.cup + U.cguid, 3);
      U.g0_g = L.f32(d + ".conv.double_conv.1.weight", U.cmid), U.g0_b = L.f32(d + ".conv.double_conv.1.bias", U.cmid);
      U.c1 = L.conv(d + ".conv.double_conv.3.weight", U.cmid, U.cmid, 3);
      U.g1_g = L.f32(d + ".conv.double_conv.4.weight", U.cmid), U.g1_b = L.f32(d + ".conv.double_conv.4.bias", U.cmid);
    }
    m->d_head_w = L.f32(p + ".head.weight", 288), m->d_head_b = L.f32(p + ".head.bias", 1);
  }
  if (L.rc != ORYON_OK) {
    set_error("oryon_backbone_finalize: %s", L.missing.c_str());
    return L.rc;
  }
  m->host.clear();
  m->finalized = true;
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// forward helpers
// ------------------------------------------------------------------------------------------------
namespace {

struct Ctx {
  oryon_handle* h;
  Backbone* m;
  cudaStream_t st;
  Arena ar;
  bool dry;   // sizing pass: allocate from a null arena, launch nothing
  int rc = ORYON_OK;
  int prec;

  SplitA split(size_t rows, int ld) {
    SplitA s;
    s.hi = ar.take<__half>(rows * ld), s.lo = ar.take<__half>(rows * ld), s.ld = ld;
    return s;
  }
  void gemm(const SplitA& A, int M, const SplitW& W, gemm::Epilogue ep, int nb0 = 1, int64_t a_b0 = 0, int64_t w_b0 = 0) {
    if (dry || rc) return;
    gemm::Problem p;
    p.M = M, p.N = W.N, p.K = W.K, p.nb0 = nb0, p.precision = W.f8x ? 2 : prec;   // f8x: A.lo was written in the LO_F8X form by its producer
    p.A.hi = A.hi, p.A.lo = A.lo, p.A.ld = A.ld, p.A.stride_b0 = a_b0;
    p.W.hi = W.hi, p.W.lo = W.lo, p.W.ld = W.ld, p.W.stride_b0 = w_b0;
    p.ep = ep;
    p.ep.alpha *= W.unscale;
    rc = gemm::launch(h, p, st);
  }
  void ln(const LnArgs& a) {
    if (dry || rc) return;
    rc = layernorm(h, a, st);
  }
  void attn(const AttnArgs& a) {
    if (dry || rc) return;
    rc = attention(h, a, st);
  }
};

gemm::Epilogue ep_f32(float* out, int ld, const float* bias, int act = gemm::ACT_NONE, const float* residual = nullptr,
                      const int32_t* row_map = nullptr) {
  gemm::Epilogue e;
  e.out32 = out, e.ld32 = ld, e.bias = bias, e.act = act, e.residual = residual, e.row_map = row_map;
  return e;
}
gemm::Epilogue ep_split(const SplitA& o, const float* bias, int act, int lo_format = gemm::LO_F16) {
  gemm::Epilogue e;
  e.out_hi = o.hi, e.out_lo = o.lo, e.ldh = o.ld, e.bias = bias, e.act = act, e.lo_format = lo_format;
  return e;
}

// CLIP residual attention blocks (clip/model.py ResidualAttentionBlock), x [n_seq*S][width] fp32 in place.
// Long sequences (the vision tower, S = 577) run attention on the tensor cores: the QKV projection writes split
// pairs, scores = alpha * Q K^T is a GEMM batched over (head, sequence) with strided operand views, a row softmax
// writes split probabilities, and P V is a second batched GEMM against V^T that writes straight into the
// concatenated-heads layout.  Short sequences (text tower, S = 77) use the shared-memory attention kernel.
void clip_blocks(Ctx& c, float* x, int n_seq, int S, int width, int heads, bool causal, const std::vector<ClipBlock>& blocks) {
  const int M = n_seq * S, d = width / heads;
  const bool tc_attn = S >= 256 && !causal && d == 64;
  if (!tc_attn && !blocks.empty() && blocks[0].out.f8x && !c.rc) {
    set_error("clip_blocks: fp8 cross-term weights need the tensor-core attention path (S >= 256, head dim 64)");
    c.rc = ORYON_ERR_INVALID_ARGUMENT;
    return;
  }
  const size_t mark = c.ar.off;
  SplitA hsp = c.split(M, width);
  SplitA att = c.split(M, width);
  SplitA hid = c.split(M, 4 * width);
  float* qkv = nullptr;
  SplitA qkvh{}, P{}, vt{};
  float* scores = nullptr;
  static const bool materialized = getenv("ORYON_ATTN_MATERIALIZED") != nullptr;   // A/B switch: scores through HBM (two batched GEMMs)
  const int ldS = round_up(S, 4), ldP = round_up(S, 128);
  const bool need_vt = materialized || !attn::v_from_qkv();   // the fused kernel reads V in place (MN-major operand)
  if (tc_attn) {
    qkvh = c.split(M, 3 * width);
    if (materialized) {
      scores = c.ar.take<float>((size_t)n_seq * heads * S * ldS);
      P = c.split((size_t)n_seq * heads * S, ldP);
    }
    if (need_vt) vt = c.split((size_t)n_seq * heads * d, ldP);
  } else {
    qkv = c.ar.take<float>((size_t)M * 3 * width);
  }
  const float scale = 1.f / sqrtf((float)d);
  for (const ClipBlock& b : blocks) {
    LnArgs l;
    l.x = x, l.ldx = width, l.C = width, l.gamma = b.ln1_g, l.beta = b.ln1_b, l.rows = M, l.out_hi = hsp.hi, l.out_lo = hsp.lo, l.ldh = width;
    l.lo_format = b.qkv.f8x;   // each producer writes the `lo` form its consumer's weights were packed for
    c.ln(l);
    if (tc_attn) {
      // precision 2: the Q and K columns leave the projection as 8-bit cross-term blocks too, for the attention kernel's Q K^T
      const bool qk8 = b.qkv.f8x && !materialized && c.prec == 3 && attn::pp_active();
      gemm::Epilogue eq = ep_split(qkvh, b.qkv_b, gemm::ACT_NONE, qk8 ? gemm::LO_QKV : gemm::LO_F16);
      eq.qkv_width = width;
      c.gemm(hsp, M, b.qkv, eq);
      if (need_vt && !c.dry && !c.rc) c.rc = transpose_v(c.h, qkvh.hi, qkvh.lo, 3 * width, 2 * width, n_seq, S, heads, d, vt.hi, vt.lo, ldP, c.st);
      if (!materialized) {
        if (!c.dry && !c.rc)
          c.rc = attn::launch(c.h, qkvh.hi, qkvh.lo, vt.hi, vt.lo, ldP, n_seq, S, heads, width, c.prec, att.hi, att.lo, width, c.st, b.out.f8x, qk8 ? 1 : 0);
      } else {
      if (!c.dry && !c.rc) {  // scores[seq][head] = scale * Q K^T
        gemm::Problem p;
        p.M = S, p.N = S, p.K = d, p.nb0 = heads, p.nb1 = n_seq, p.precision = c.prec;
        p.A.hi = qkvh.hi, p.A.lo = qkvh.lo, p.A.ld = 3 * width, p.A.stride_b0 = d, p.A.stride_b1 = (int64_t)S * 3 * width;
        p.W.hi = qkvh.hi + width, p.W.lo = qkvh.lo + width, p.W.ld = 3 * width, p.W.stride_b0 = d, p.W.stride_b1 = (int64_t)S * 3 * width;
        p.ep.alpha = scale, p.ep.out32 = scores, p.ep.ld32 = ldS, p.ep.out_b0 = (int64_t)S * ldS, p.ep.out_b1 = (int64_t)heads * S * ldS;
        c.rc = gemm::launch(c.h, p, c.st);
      }
      if (!c.dry && !c.rc) c.rc = softmax_split(c.h, scores, (int64_t)n_seq * heads * S, S, ldS, P.hi, P.lo, ldP, c.st);
      if (!c.dry && !c.rc) {  // out[seq][s][head*d + :] = P V
        gemm::Problem p;
        p.M = S, p.N = d, p.K = S, p.nb0 = heads, p.nb1 = n_seq, p.precision = c.prec;
        p.A.hi = P.hi, p.A.lo = P.lo, p.A.ld = ldP, p.A.stride_b0 = (int64_t)S * ldP, p.A.stride_b1 = (int64_t)heads * S * ldP;
        p.W.hi = vt.hi, p.W.lo = vt.lo, p.W.ld = ldP, p.W.stride_b0 = (int64_t)d * ldP, p.W.stride_b1 = (int64_t)heads * d * ldP;
        p.ep.out_hi = att.hi, p.ep.out_lo = att.lo, p.ep.ldh = width, p.ep.outh_b0 = d, p.ep.outh_b1 = (int64_t)S * width;
        p.ep.lo_format = b.out.f8x;
        c.rc = gemm::launch(c.h, p, c.st);
      }
      }
    } else {
      c.gemm(hsp, M, b.qkv, ep_f32(qkv, 3 * width, b.qkv_b));
      AttnArgs a;
      a.qkv = qkv, a.ld = 3 * width, a.off_k = width, a.off_v = 2 * width, a.n_seq = n_seq, a.S = S, a.heads = heads, a.d = d;
      a.scale = scale, a.causal = causal ? 1 : 0, a.out_hi = att.hi, a.out_lo = att.lo, a.ldh = width;
      c.attn(a);
    }
    c.gemm(att, M, b.out, ep_f32(x, width, b.out_b, gemm::ACT_NONE, x));
    l.gamma = b.ln2_g, l.beta = b.ln2_b, l.lo_format = b.fc.f8x;
    c.ln(l);
    c.gemm(hsp, M, b.fc, ep_split(hid, b.fc_b, gemm::ACT_QUICKGELU, b.proj.f8x));
    c.gemm(hid, M, b.proj, ep_f32(x, width, b.proj_b, gemm::ACT_NONE, x));
  }
  c.ar.off = mark;  // scratch of the tower is reusable afterwards
}

// One Swin block on x [N*H*W][dim] (in place).  guid != null: fusion variant (q,k from cat(norm1(x), guid)).
void swin_block(Ctx& c, float* x, int N, const Stage& S, const SwinBlock& B, bool shifted, const int32_t* map, const float* guid, int mlp_act) {
  const int dim = S.dim, ws2 = S.ws * S.ws, Mw = N * S.nW * ws2, M = N * S.H * S.W;
  const int kin = guid ? 2 * dim : dim;
  const size_t mark = c.ar.off;
  SplitA a_in = c.split(Mw, kin);
  float* qkv = c.ar.take<float>((size_t)Mw * 3 * dim);
  SplitA att = c.split(Mw, dim);
  LnArgs l;
  l.x = x, l.ldx = dim, l.C = dim, l.gamma = B.n1_g, l.beta = B.n1_b, l.rows = Mw, l.row_map = map, l.cat = guid, l.cat_C = guid ? dim : 0;
  l.out_hi = a_in.hi, l.out_lo = a_in.lo, l.ldh = kin;
  c.ln(l);
  c.gemm(a_in, Mw, B.qkv, ep_f32(qkv, 3 * dim, B.qkv_b));
  AttnArgs a;
  a.qkv = qkv, a.ld = 3 * dim, a.off_k = dim, a.off_v = 2 * dim, a.n_seq = N * S.nW, a.S = ws2, a.heads = S.heads, a.d = dim / S.heads;
  a.scale = 1.f / sqrtf((float)(dim / S.heads)), a.bias = B.bias, a.bias_t = B.bias_t, a.mask = shifted ? S.mask : nullptr, a.n_win = S.nW;
  a.out_hi = att.hi, a.out_lo = att.lo, a.ldh = dim;
  c.attn(a);
  c.gemm(att, Mw, B.proj, ep_f32(x, dim, B.proj_b, gemm::ACT_NONE, x, map));
  // MLP
  SplitA hsp = c.split(M, dim), hid = c.split(M, 4 * dim);
  LnArgs l2;
  l2.x = x, l2.ldx = dim, l2.C = dim, l2.gamma = B.n2_g, l2.beta = B.n2_b, l2.rows = M, l2.out_hi = hsp.hi, l2.out_lo = hsp.lo, l2.ldh = dim;
  c.ln(l2);
  c.gemm(hsp, M, B.fc1, ep_split(hid, B.fc1_b, mlp_act));
  c.gemm(hid, M, B.fc2, ep_f32(x, dim, B.fc2_b, gemm::ACT_NONE, x));
  c.ar.off = mark;
}

// 3x3 / 7x7 convolution; returns fp32 NHWC [n*H*W][cout].  Implicit GEMM: the A operand is gathered from the activations by
// producer warps inside the GEMM kernel (gemm::ConvGather); the im2col matrix (up to 1.5 GB per layer at 192x192) is neither
// written nor read.  ORYON_CONV_IM2COL=1 (A/B switch) or channel counts that are not multiples of 8 take im2col + GEMM.
float* conv_nhwc(Ctx& c, const Im2colArgs& src, const SplitW& W, const float* bias, int act) {
  const size_t rows = (size_t)src.n * src.H * src.W;
  float* out = c.ar.take<float>(rows * W.N);
  static const bool force_im2col = getenv("ORYON_CONV_IM2COL") != nullptr;
  if (!force_im2col && src.C0 % 8 == 0 && src.C1 % 8 == 0 && W.K == src.k * src.k * (src.C0 + src.C1)) {
    if (!c.dry && !c.rc) {
      gemm::ConvGather g;
      g.src0 = src.src0, g.C0 = src.C0, g.shuffle0 = src.shuffle0, g.src1 = src.src1, g.C1 = src.C1;
      g.n = src.n, g.H = src.H, g.W = src.W, g.k = src.k;
      gemm::Problem p;
      p.M = (int)rows, p.N = W.N, p.K = W.K, p.precision = c.prec;
      p.W.hi = W.hi, p.W.lo = W.lo, p.W.ld = W.ld;
      p.ep = ep_f32(out, W.N, bias, act);
      p.gather = &g;
      c.rc = gemm::launch(c.h, p, c.st);
    }
    return out;
  }
  const size_t mark = c.ar.off;
  SplitA A = c.split(rows, W.ld);
  if (!c.dry && !c.rc) {
    Im2colArgs a = src;
    a.hi = A.hi, a.lo = A.lo, a.ld = W.ld;
    c.rc = im2col(c.h, a, c.st);
  }
  c.gemm(A, (int)rows, W, ep_f32(out, W.N, bias, act));
  c.ar.off = mark;
  return out;
}

const std::vector<int32_t*>& window_maps(Ctx& c, int N) {
  // [0,1] guidance stage 0 (plain, shifted)  [2,3] guidance stage 1  [4,5] fusion  [6] patch-token rows of the CLIP sequence
  auto it = c.m->maps.find(N);
  if (it != c.m->maps.end()) return it->second;
  std::vector<int32_t*> v;
  const Stage* stages[3] = {&c.m->s_stage[0], &c.m->s_stage[1], &c.m->f_stage};
  for (const Stage* s : stages)
    for (int sh = 0; sh < 2; ++sh) {
      const std::vector<int32_t> hm = window_map(N, *s, sh ? s->shift : 0);
      int32_t* d = nullptr;
      if (cudaMalloc(&d, hm.size() * 4) == cudaSuccess) {
        cudaMemcpyAsync(d, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice, c.st);
        cudaStreamSynchronize(c.st);
      } else if (!c.rc) {
        c.rc = ORYON_ERR_OUT_OF_MEMORY;
      }
      v.push_back(d);
    }
  {
    std::vector<int32_t> hr((size_t)N * kVisTok);  // x[:, 1:, :] (vlm.py:56)
    for (int n = 0; n < N; ++n)
      for (int t = 0; t < kVisTok; ++t) hr[(size_t)n * kVisTok + t] = n * (kVisTok + 1) + 1 + t;
    int32_t* d = nullptr;
    if (cudaMalloc(&d, hr.size() * 4) == cudaSuccess) {
      cudaMemcpyAsync(d, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice, c.st);
      cudaStreamSynchronize(c.st);
    } else if (!c.rc) {
      c.rc = ORYON_ERR_OUT_OF_MEMORY;
    }
    v.push_back(d);
  }
  return c.m->maps.emplace(N, v).first->second;
}

const float kClipMean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, kClipStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};
const float kImnMean[3] = {0.485f, 0.456f, 0.406f}, kImnStd[3] = {0.229f, 0.224f, 0.225f};

struct Outputs {
  float *feat_a, *feat_q, *mask_a, *mask_q;
  const oryon_backbone_debug* dbg;
};

// The network for Bc pairs starting at pair b0 of the call.
void forward_chunk(Ctx& c, const float* rgb_a, const float* rgb_q, int b0, int Bc, const float* text, const Outputs& o) {
  Backbone* m = c.m;
  const int N = 2 * Bc;
  const size_t img_el = (size_t)3 * 224 * 224;
  c.ar.off = 0;
  // images of the chunk, anchors first then queries, contiguous
  float* rgb = c.ar.take<float>((size_t)N * img_el);
  if (!c.dry && !c.rc) {
    cudaMemcpyAsync(rgb, rgb_a + (size_t)b0 * img_el, (size_t)Bc * img_el * 4, cudaMemcpyDeviceToDevice, c.st);
    cudaMemcpyAsync(rgb + (size_t)Bc * img_el, rgb_q + (size_t)b0 * img_el, (size_t)Bc * img_el * 4, cudaMemcpyDeviceToDevice, c.st);
  }
  const std::vector<int32_t*>* maps = c.dry ? nullptr : &window_maps(c, N);
  auto map = [&](int i) -> const int32_t* { return maps ? (*maps)[i] : nullptr; };

  // ---------------- CLIP vision tower (vlm.py:43-61) ----------------
  const int Mv = N * (kVisTok + 1);
  float* xv = c.ar.take<float>((size_t)Mv * kVisW);
  SplitA vtok = c.split((size_t)N * kVisTok, kVisW);  // ln_post(patch tokens)
  {
    const size_t mark = c.ar.off;
    SplitA pa = c.split((size_t)N * kVisTok, m->v_patch.ld);
    float* pe = c.ar.take<float>((size_t)N * kVisTok * kVisW);
    if (!c.dry && !c.rc) c.rc = resize_patch(c.h, rgb, N, 224, 336, 14, 0, kClipMean, kClipStd, pa.hi, pa.lo, pa.ld, c.st);
    c.gemm(pa, N * kVisTok, m->v_patch, ep_f32(pe, kVisW, nullptr));
    if (!c.dry && !c.rc) c.rc = clip_embed_ln(c.h, pe, m->v_cls, m->v_pos, m->v_lnpre_g, m->v_lnpre_b, N, kVisTok, kVisW, xv, c.st);
    c.ar.off = mark;
  }
  clip_blocks(c, xv, N, kVisTok + 1, kVisW, kVisHeads, false, m->v_blocks);
  {  // ln_post on the patch tokens only (x[:, 1:, :])
    LnArgs l;
    l.x = xv, l.ldx = kVisW, l.C = kVisW, l.gamma = m->v_lnpost_g, l.beta = m->v_lnpost_b, l.rows = N * kVisTok, l.row_map = map(6);
    l.out_hi = vtok.hi, l.out_lo = vtok.lo, l.ldh = kVisW;
    float* dbg32 = nullptr;
    if (o.dbg && o.dbg->clip_tokens) dbg32 = c.ar.take<float>((size_t)N * kVisTok * kVisW), l.out32 = dbg32, l.ld32 = kVisW;
    c.ln(l);
    if (dbg32 && !c.dry && !c.rc) {  // [N][1024][24][24], anchors at [b0..), queries at [B + b0..)
      c.rc = nhwc_to_nchw(c.h, dbg32, Bc, kVisTok, kVisW, o.dbg->clip_tokens + (size_t)b0 * kVisTok * kVisW, c.st);
      if (!c.rc)
        c.rc = nhwc_to_nchw(c.h, dbg32 + (size_t)Bc * kVisTok * kVisW, Bc, kVisTok, kVisW,
                            o.dbg->clip_tokens + (size_t)(o.dbg->B + b0) * kVisTok * kVisW, c.st);
    }
  }

  // ---------------- guidance backbone (net.py:60-75) ----------------
  float* g3 = c.ar.take<float>((size_t)N * 9216 * 128);   // [N][96][96][128]
  float* g2 = c.ar.take<float>((size_t)N * 2304 * 256);   // [N][48][48][256]
  float* g1 = c.ar.take<float>((size_t)N * 576 * 512);    // [N][24][24][512]
  {
    const size_t mark = c.ar.off;
    SplitA pa = c.split((size_t)N * 9216, m->s_patch.ld);
    if (!c.dry && !c.rc) c.rc = resize_patch(c.h, rgb, N, 224, 384, 4, 1, kImnMean, kImnStd, pa.hi, pa.lo, pa.ld, c.st);
    c.gemm(pa, N * 9216, m->s_patch, ep_f32(g3, 128, m->s_patch_b));
    LnArgs l;
    l.x = g3, l.ldx = 128, l.C = 128, l.gamma = m->s_pn_g, l.beta = m->s_pn_b, l.rows = N * 9216, l.out32 = g3, l.ld32 = 128;
    c.ln(l);
    c.ar.off = mark;
    swin_block(c, g3, N, m->s_stage[0], m->s_blk[0][0], false, map(0), nullptr, gemm::ACT_GELU);
    swin_block(c, g3, N, m->s_stage[0], m->s_blk[0][1], true, map(1), nullptr, gemm::ACT_GELU);
    SplitA mg = c.split((size_t)N * 2304, 512);
    if (!c.dry && !c.rc) c.rc = patch_merge_ln(c.h, g3, N, 96, 96, 128, m->s_mg_g[0], m->s_mg_b[0], mg.hi, mg.lo, c.st);
    c.gemm(mg, N * 2304, m->s_merge[0], ep_f32(g2, 256, nullptr));
    c.ar.off = mark;
    // guidance2 is the output of features.2 (before stage features.3 runs): keep it, run the stage on a copy
    float* x2 = c.ar.take<float>((size_t)N * 2304 * 256);
    if (!c.dry && !c.rc) cudaMemcpyAsync(x2, g2, (size_t)N * 2304 * 256 * 4, cudaMemcpyDeviceToDevice, c.st);
    swin_block(c, x2, N, m->s_stage[1], m->s_blk[1][0], false, map(2), nullptr, gemm::ACT_GELU);
    swin_block(c, x2, N, m->s_stage[1], m->s_blk[1][1], true, map(3), nullptr, gemm::ACT_GELU);
    SplitA mg2 = c.split((size_t)N * 576, 1024);
    if (!c.dry && !c.rc) c.rc = patch_merge_ln(c.h, x2, N, 48, 48, 256, m->s_mg_g[1], m->s_mg_b[1], mg2.hi, mg2.lo, c.st);
    c.gemm(mg2, N * 576, m->s_merge[1], ep_f32(g1, 512, nullptr));
    c.ar.off = mark;
    if (o.dbg && !c.dry && !c.rc) {
      float* dst[3] = {o.dbg->guid1, o.dbg->guid2, o.dbg->guid3};
      const float* src[3] = {g1, g2, g3};
      const int hw[3] = {576, 2304, 9216}, ch[3] = {512, 256, 128};
      for (int i = 0; i < 3 && !c.rc; ++i)
        if (dst[i]) {
          const size_t per = (size_t)hw[i] * ch[i];
          c.rc = nhwc_to_nchw(c.h, src[i], Bc, hw[i], ch[i], dst[i] + (size_t)b0 * per, c.st);
          if (!c.rc) c.rc = nhwc_to_nchw(c.h, src[i] + (size_t)Bc * per, Bc, hw[i], ch[i], dst[i] + (size_t)(o.dbg->B + b0) * per, c.st);
        }
    }
  }

  // ---------------- fusion (fusion.py:602-625) ----------------
  float* xf = nullptr;  // [N][576][128]
  {
    float* proj = c.ar.take<float>((size_t)N * kVisTok * 768);
    c.gemm(vtok, N * kVisTok, m->f_clipconv, ep_f32(proj, 768, m->f_clipconv_b));
    SplitA pn = c.split((size_t)N * kVisTok, 768);
    SplitA tn = c.split((size_t)Bc * kPrompts, 768);
    float* corr = c.ar.take<float>((size_t)N * kVisTok * kPrompts);   // [N][24*24][80]
    const float* text_c = text + (size_t)b0 * kPrompts * 768;
    if (!c.dry && !c.rc) c.rc = l2norm_split(c.h, proj, N * kVisTok, 768, pn.hi, pn.lo, 768, c.st);
    if (!c.dry && !c.rc) c.rc = l2norm_split(c.h, text_c, Bc * kPrompts, 768, tn.hi, tn.lo, 768, c.st);
    SplitW tw;
    tw.hi = tn.hi, tw.lo = tn.lo, tw.N = kPrompts, tw.K = 768, tw.ld = 768;
    for (int half = 0; half < 2; ++half) {  // anchors, then queries: both use the pair's prompt embeddings
      SplitA a = pn;
      a.hi += (size_t)half * Bc * kVisTok * 768, a.lo += (size_t)half * Bc * kVisTok * 768;
      gemm::Epilogue e = ep_f32(corr + (size_t)half * Bc * kVisTok * kPrompts, kPrompts, nullptr);
      e.out_b0 = (int64_t)kVisTok * kPrompts;
      c.gemm(a, kVisTok, tw, e, Bc, (int64_t)kVisTok * 768, (int64_t)kPrompts * 768);
    }
    Im2colArgs ic;
    ic.src0 = corr, ic.C0 = kPrompts, ic.n = N, ic.H = 24, ic.W = 24, ic.k = 7;
    xf = conv_nhwc(c, ic, m->f_conv1, m->f_conv1_b, gemm::ACT_NONE);
    Im2colArgs ig;
    ig.src0 = g1, ig.C0 = 512, ig.n = N, ig.H = 24, ig.W = 24, ig.k = 3;
    float* pg = conv_nhwc(c, ig, m->f_guid, m->f_guid_b, gemm::ACT_RELU);
    float* tg = c.ar.take<float>((size_t)Bc * 128);
    if (!c.dry && !c.rc) c.rc = text_guidance(c.h, text_c, Bc, kPrompts, 768, m->f_text_w, m->f_text_b, 128, tg, c.st);
    float* gln = c.ar.take<float>((size_t)N * 576 * 128);
    for (int li = 0; li < 2; ++li) {
      LnArgs l;
      l.x = pg, l.ldx = 128, l.C = 128, l.gamma = m->f_gn_g[li], l.beta = m->f_gn_b[li], l.rows = N * 576, l.out32 = gln, l.ld32 = 128;
      c.ln(l);
      swin_block(c, xf, N, m->f_stage, m->f_blk[li][0], false, map(4), gln, gemm::ACT_GELU);
      swin_block(c, xf, N, m->f_stage, m->f_blk[li][1], true, map(5), gln, gemm::ACT_GELU);
      if (!c.dry && !c.rc) c.rc = class_transformer(c.h, xf, tg, N, Bc, m->f_ct[li], c.st);
    }
    if (o.dbg && o.dbg->fusion && !c.dry && !c.rc) {
      const size_t per = (size_t)576 * 128;
      c.rc = nhwc_to_nchw(c.h, xf, Bc, 576, 128, o.dbg->fusion + (size_t)b0 * per, c.st);
      if (!c.rc) c.rc = nhwc_to_nchw(c.h, xf + (size_t)Bc * per, Bc, 576, 128, o.dbg->fusion + (size_t)(o.dbg->B + b0) * per, c.st);
    }
  }

  // ---------------- decoder (decoder.py:82-108) ----------------
  {
    Im2colArgs i0;
    i0.src0 = g2, i0.C0 = 256, i0.n = N, i0.H = 48, i0.W = 48, i0.k = 3;
    float* pg0 = conv_nhwc(c, i0, m->d_g0, m->d_g0_b, gemm::ACT_RELU);    // [N][48][48][32]
    Im2colArgs i1;
    i1.src0 = g3, i1.C0 = 128, i1.n = N, i1.H = 96, i1.W = 96, i1.k = 3;
    float* pg1 = conv_nhwc(c, i1, m->d_g1, m->d_g1_b, gemm::ACT_RELU);    // [N][96][96][16]
    double* gstats = c.ar.take<double>((size_t)N * 8);
    float* cur = xf;
    int H = 24;
    const float* guid[3] = {pg0, pg1, nullptr};
    for (int i = 0; i < 3; ++i) {
      const UpW& U = m->d_up[i];
      const size_t rows = (size_t)N * H * H;
      // ConvTranspose2d(k=2, s=2): one GEMM producing the four sub-pixels of every input pixel
      SplitA a = c.split(rows, round_up(U.cin, 64));
      if (!c.dry && !c.rc) c.rc = gemm::split_rows(c.h, cur, U.cin, (int)rows, U.cin, a.hi, a.lo, a.ld, c.st);
      float* up = c.ar.take<float>(rows * 4 * U.cup);
      c.gemm(a, (int)rows, U.up, ep_f32(up, 4 * U.cup, U.up_b4));
      H *= 2;
      Im2colArgs ia;
      ia.src0 = up, ia.C0 = U.cup, ia.shuffle0 = 1, ia.src1 = guid[i], ia.C1 = U.cguid, ia.n = N, ia.H = H, ia.W = H, ia.k = 3;
      float* y = conv_nhwc(c, ia, U.c0, nullptr, gemm::ACT_NONE);
      if (!c.dry && !c.rc) c.rc = groupnorm_relu(c.h, y, N, H * H, U.cmid, U.g0_g, U.g0_b, gstats, c.st);
      Im2colArgs ib;
      ib.src0 = y, ib.C0 = U.cmid, ib.n = N, ib.H = H, ib.W = H, ib.k = 3;
      float* z = conv_nhwc(c, ib, U.c1, nullptr, gemm::ACT_NONE);
      if (!c.dry && !c.rc) c.rc = groupnorm_relu(c.h, z, N, H * H, U.cmid, U.g1_g, U.g1_b, gstats, c.st);
      cur = z;
    }
    // head + outputs: anchors are images [0, Bc), queries [Bc, 2Bc) of the chunk
    const size_t fm = (size_t)32 * 192 * 192, lg = (size_t)192 * 192;
    if (!c.dry && !c.rc)
      c.rc = decoder_head(c.h, cur, Bc, 192, 192, m->d_head_w, m->d_head_b, o.mask_a + (size_t)b0 * lg, o.feat_a + (size_t)b0 * fm, c.st);
    if (!c.dry && !c.rc)
      c.rc = decoder_head(c.h, cur + (size_t)Bc * 192 * 192 * 32, Bc, 192, 192, m->d_head_w, m->d_head_b, o.mask_q + (size_t)b0 * lg,
                          o.feat_q + (size_t)b0 * fm, c.st);
  }
}

}  // namespace

int backbone_forward(oryon_handle* h, const float* rgb_a, const float* rgb_q, int B, const float* text_emb, float* feat_a, float* feat_q,
                     float* mask_a, float* mask_q, const oryon_backbone_debug* dbg, cudaStream_t st) {
  ORYON_REQUIRE(h && rgb_a && rgb_q && text_emb && feat_a && feat_q && mask_a && mask_q && B > 0, "oryon_backbone_forward: bad argument");
  Backbone* m = model_of(h);
  if (!m || !m->finalized) {
    set_error("oryon_backbone_forward: weights not loaded / finalized");
    return ORYON_ERR_NOT_LOADED;
  }
  ORYON_REQUIRE(!dbg || dbg->B == B, "oryon_backbone_forward: debug.B must equal B");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  const int chunk = std::max(1, std::min(B, m->cfg.max_pairs_per_pass > 0 ? m->cfg.max_pairs_per_pass : 16));
  Outputs o{feat_a, feat_q, mask_a, mask_q, dbg};
  Ctx c{h, m, st, Arena(), true};
  c.prec = m->cfg.precision == 2 ? 3 : m->cfg.precision;   // 2: three products everywhere but the layers whose weights are packed f8x
  forward_chunk(c, rgb_a, rgb_q, 0, chunk, text_emb, o);   // sizing pass
  int rc;
  if ((rc = m->arena.reserve(c.ar.peak + 1024, st))) return rc;
  c.dry = false;
  c.ar.base = m->arena.as<char>();
  for (int b0 = 0; b0 < B && !c.rc; b0 += chunk) forward_chunk(c, rgb_a, rgb_q, b0, std::min(chunk, B - b0), text_emb, o);
  if (c.rc == ORYON_OK) ORYON_CUDA_CHECK(cudaGetLastError());
  return c.rc;
}

int text_forward(oryon_handle* h, const int32_t* tokens, int n, float* out, cudaStream_t st) {
  ORYON_REQUIRE(h && tokens && out && n > 0, "oryon_text_forward: bad argument");
  Backbone* m = model_of(h);
  if (!m || !m->finalized || m->t_blocks.empty()) {
    set_error("oryon_text_forward: text tower not loaded");
    return ORYON_ERR_NOT_LOADED;
  }
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  const int chunk = std::min(n, 640);
  Ctx c{h, m, st, Arena(), true};
  c.prec = m->cfg.precision == 2 ? 3 : m->cfg.precision;   // 2: three products everywhere but the layers whose weights are packed f8x
  auto run = [&](int s0, int ns) {
    c.ar.off = 0;
    float* x = c.ar.take<float>((size_t)ns * kTxtL * kTxtW);
    int32_t* rows = c.ar.take<int32_t>(ns);
    SplitA fin = c.split(ns, kTxtW);
    if (!c.dry && !c.rc) c.rc = text_embed(h, tokens + (size_t)s0 * kTxtL, m->t_tok, m->t_pos, ns, kTxtL, kTxtW, kVocab, x, st);
    clip_blocks(c, x, ns, kTxtL, kTxtW, kTxtHeads, true, m->t_blocks);
    if (!c.dry && !c.rc) c.rc = eot_rows(h, tokens + (size_t)s0 * kTxtL, ns, kTxtL, rows, st);
    LnArgs l;
    l.x = x, l.ldx = kTxtW, l.C = kTxtW, l.gamma = m->t_lnf_g, l.beta = m->t_lnf_b, l.rows = ns, l.row_map = rows;
    l.out_hi = fin.hi, l.out_lo = fin.lo, l.ldh = kTxtW;
    c.ln(l);
    c.gemm(fin, ns, m->t_proj, ep_f32(out + (size_t)s0 * kTxtW, kTxtW, nullptr));
  };
  run(0, chunk);
  int rc;
  // the text tower shares the arena with the image path (calls are stream ordered)
  if ((rc = m->arena.reserve(c.ar.peak + 1024, st))) return rc;
  c.dry = false;
  c.ar.base = m->arena.as<char>();
  for (int s0 = 0; s0 < n && !c.rc; s0 += chunk) run(s0, std::min(chunk, n - s0));
  return c.rc;
}

}  // namespace net
}  // namespace oryon
