// engine.cu -- host-side execution plan for the image_transformer_v2 denoiser
// (reference: k_diffusion/models/image_transformer_v2.py:667-762 and layers.py:88-90).
//
// The engine owns no activations: the caller passes one workspace; the plan carves it.  Weights are
// borrowed device pointers keyed by the reference state-dict names; kdb_model_finalize builds the
// derived tables (bf16 copies, concatenated AdaRMSNorm projection, position grids).
#include <cmath>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "model_kernels.cuh"
#include "tc_kernels.cuh"

namespace kdb {

struct TensorRef {
  const float* p = nullptr;
  std::vector<int64_t> shape;
};

struct LayerPlan {
  std::string prefix;
  int level = 0, attn_type = 0, attn_param = 0, shift = 0;
  int C = 0, dff = 0, nh = 0, e = 0;
  int ada_attn = -1, ada_ff = -1;   // offsets into a conditioning row
  const float *attn_norm_w = nullptr, *qkv_w = nullptr, *scale = nullptr, *freqs = nullptr, *out_w = nullptr;
  const float *ff_norm_w = nullptr, *up_w = nullptr, *down_w = nullptr;
  bf16 *qkv_wb = nullptr, *out_wb = nullptr, *up_wb = nullptr, *down_wb = nullptr;
  bf16* up_wb_il = nullptr;          // up_proj rows interleaved (value/gate) for the fused GEGLU epilogue
  int exec_index = 0;                // position in execution order (indexes PosTables::rope)
  bool bounded = false;              // every scale[h] in (0, KDB_ATTN_MAX_BOUND]: the scale is the attention kernels' fixed softmax shift
  bf16 *qkv_wf = nullptr, *up_wf = nullptr;   // per-evaluation copies with the AdaRMSNorm channel scale folded in (fused norm)
};

struct PosTables {
  std::vector<float*> pos;          // per level: [T_l, 2] (y, x)
  std::vector<float2*> rope;        // per layer (execution order): [T_l, nh, 16] (cos, sin) of the RoPE angles, or nullptr
};

}  // namespace kdb

using namespace kdb;

struct KdbModel {
  KdbModelConfig cfg{};
  std::unordered_map<std::string, TensorRef> tensors;
  bool finalized = false;
  std::vector<std::vector<LayerPlan>> down, up;
  std::vector<LayerPlan> mid;
  std::vector<const float*> merge_w, split_w, split_fac;
  std::vector<bf16*> merge_wb, split_wb;
  std::vector<void*> owned;
  float* ada_cat = nullptr;
  FoldDesc* fold_descs = nullptr;   // device table for launch_fold_norm_weights
  int n_fold = 0;
  bool fuse_norm = true;
  bool ss_valid = false;            // ws.rowss describes the current residual stream (set by the GEMM that produced it)
  bf16* patch_out_wb = nullptr;     // patch_out.proj.weight zero-padded to 64 rows (tensor-core patch-out)
  bf16* patch_out_wf = nullptr;     // the same with out_norm.scale folded in (fused out_norm)
  bf16* patch_in_wb = nullptr;      // patch_in.proj.weight, columns permuted to (c, nh, nw) and padded to 64 (tensor-core patch-in)
  int ada_total = 0;
  CondWeights cw{};
  std::map<std::pair<int, int>, PosTables> pos_cache;
  // tap
  std::string tap_name;
  float* tap_out = nullptr;
  int64_t tap_cap = 0, tap_count = 0;
  int layer_counter = 0;
  int n_layers = 0;
};

namespace {

int get(KdbModel* m, const std::string& key, std::initializer_list<int64_t> shape, const float** out) {
  auto it = m->tensors.find(key);
  if (it == m->tensors.end()) {
    set_error("missing state-dict entry '%s'", key.c_str());
    return KDB_ERR_MISSING_KEY;
  }
  const std::vector<int64_t> want(shape);
  if (it->second.shape != want) {
    std::string got, exp;
    for (auto v : it->second.shape) got += std::to_string(v) + ",";
    for (auto v : want) exp += std::to_string(v) + ",";
    set_error("shape mismatch for '%s': got [%s] expected [%s]", key.c_str(), got.c_str(), exp.c_str());
    return KDB_ERR_BAD_SHAPE;
  }
  *out = it->second.p;
  return 0;
}

#define GET(key, out, ...)                                        \
  do {                                                            \
    int rc__ = get(m, (key), {__VA_ARGS__}, (out));               \
    if (rc__) return rc__;                                        \
  } while (0)

template <typename T>
int dev_alloc(KdbModel* m, T** p, size_t count) {
  void* q = nullptr;
  KDB_CUDA(cudaMalloc(&q, count * sizeof(T) + 1024));
  m->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

void free_owned(KdbModel* m) {
  for (void* p : m->owned) cudaFree(p);
  m->owned.clear();
  m->pos_cache.clear();
}

int make_bf16(KdbModel* m, const float* src, int64_t n, bf16** dst, cudaStream_t st) {
  int rc = dev_alloc(m, dst, (size_t)n);
  if (rc) return rc;
  return launch_f32_to_bf16(src, *dst, n, st);
}

// rows of up_proj [2F, C] reordered so that every 16-row group holds 8 value rows followed by the
// 8 matching gate rows: lets a tensor-core epilogue that owns >= 16 consecutive columns apply GEGLU.
__global__ void interleave_geglu_rows_kernel(const float* __restrict__ w, bf16* __restrict__ out, int F, int C) {
  const int64_t total = (int64_t)2 * F * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const int64_t g = r / 16;
    const int j = (int)(r - g * 16);
    const int64_t src_row = (j < 8) ? (g * 8 + j) : ((int64_t)F + g * 8 + (j - 8));
    out[i] = __float2bfloat16_rn(w[src_row * C + c]);
  }
}

int plan_layer(KdbModel* m, LayerPlan& L, const std::string& prefix, int level, int index, int* ada_off, cudaStream_t st) {
  const KdbModelConfig& c = m->cfg;
  const int mw = c.mapping_width;
  L.prefix = prefix;
  L.level = level;
  L.C = c.width[level];
  L.dff = c.d_ff[level];
  L.attn_type = c.attn_type[level];
  L.attn_param = c.attn_param[level];
  if (L.attn_type != KDB_ATTN_NONE) {
    L.e = c.d_head[level];
    KDB_REQUIRE(L.e > 0 && L.C % L.e == 0, KDB_ERR_BAD_SHAPE, "level %d: width %d not divisible by d_head %d", level, L.C, L.e);
    L.nh = L.C / L.e;
    // image_transformer_v2.py:523 -- odd layer index => shift by half a window
    L.shift = (L.attn_type == KDB_ATTN_SHIFTED_WINDOW && (index % 2 == 1)) ? L.attn_param / 2 : 0;
    const std::string a = prefix + "self_attn.";
    GET(a + "norm.linear.weight", &L.attn_norm_w, L.C, mw);
    GET(a + "qkv_proj.weight", &L.qkv_w, 3 * L.C, L.C);
    GET(a + "scale", &L.scale, L.nh);
    GET(a + "pos_emb.freqs", &L.freqs, L.nh, L.e / 8);
    GET(a + "out_proj.weight", &L.out_w, L.C, L.C);
    L.ada_attn = *ada_off;
    *ada_off += L.C;
    {   // |q . k| <= scale_h after the cosine-similarity normalisation: usable as a fixed softmax shift while exp(-2 scale) stays normal
      std::vector<float> hs((size_t)L.nh);
      KDB_CUDA(cudaMemcpyAsync(hs.data(), L.scale, sizeof(float) * L.nh, cudaMemcpyDeviceToHost, st));
      KDB_CUDA(cudaStreamSynchronize(st));
      L.bounded = true;
      for (float v : hs) L.bounded = L.bounded && v > 0.f && v <= KDB_ATTN_MAX_BOUND;
    }
    int rc;
    if ((rc = make_bf16(m, L.qkv_w, 3LL * L.C * L.C, &L.qkv_wb, st))) return rc;
    if ((rc = make_bf16(m, L.out_w, (int64_t)L.C * L.C, &L.out_wb, st))) return rc;
    if ((rc = dev_alloc(m, &L.qkv_wf, (size_t)3 * L.C * L.C))) return rc;
  }
  const std::string f = prefix + "ff.";
  GET(f + "norm.linear.weight", &L.ff_norm_w, L.C, mw);
  GET(f + "up_proj.weight", &L.up_w, 2 * L.dff, L.C);
  GET(f + "down_proj.weight", &L.down_w, L.C, L.dff);
  L.ada_ff = *ada_off;
  *ada_off += L.C;
  int rc;
  if ((rc = make_bf16(m, L.up_w, 2LL * L.dff * L.C, &L.up_wb, st))) return rc;
  if ((rc = make_bf16(m, L.down_w, (int64_t)L.C * L.dff, &L.down_wb, st))) return rc;
  if (L.dff % 8 == 0) {
    if ((rc = dev_alloc(m, &L.up_wb_il, (size_t)2 * L.dff * L.C))) return rc;
    interleave_geglu_rows_kernel<<<kNumSMs * 4, 256, 0, st>>>(L.up_w, L.up_wb_il, L.dff, L.C);
    KDB_LAUNCH_CHECK(F_CONVERT, st);
    if ((rc = dev_alloc(m, &L.up_wf, (size_t)2 * L.dff * L.C))) return rc;
  }
  return 0;
}

int ensure_pos(KdbModel* m, int h0, int w0, cudaStream_t st, PosTables** out) {
  auto key = std::make_pair(h0, w0);
  auto it = m->pos_cache.find(key);
  if (it != m->pos_cache.end()) {
    *out = &it->second;
    return 0;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  KDB_CUDA(cudaStreamIsCapturing(st, &cs));
  KDB_REQUIRE(cs == cudaStreamCaptureStatusNone, KDB_ERR_UNSUPPORTED,
              "first forward for a new token grid (%dx%d) must run outside CUDA-graph capture", h0, w0);
  // axial_rope.py:31-68: cell centres in [-1,1] (short side scaled by the aspect ratio), (y, x) order
  const double ar = (double)w0 / (double)h0;
  const double ys = ar > 1.0 ? 1.0 / ar : 1.0, xs = ar < 1.0 ? ar : 1.0;
  std::vector<double> cur((size_t)h0 * w0 * 2);
  for (int i = 0; i < h0; ++i)
    for (int j = 0; j < w0; ++j) {
      cur[((size_t)i * w0 + j) * 2 + 0] = ((2.0 * i + 1.0) / h0 - 1.0) * ys;
      cur[((size_t)i * w0 + j) * 2 + 1] = ((2.0 * j + 1.0) / w0 - 1.0) * xs;
    }
  PosTables pt;
  int h = h0, w = w0;
  for (int l = 0; l < m->cfg.n_levels; ++l) {
    std::vector<float> f(cur.begin(), cur.end());
    float* d = nullptr;
    int rc = dev_alloc(m, &d, f.size());
    if (rc) return rc;
    KDB_CUDA(cudaMemcpyAsync(d, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    KDB_CUDA(cudaStreamSynchronize(st));
    pt.pos.push_back(d);
    if (l + 1 < m->cfg.n_levels) {   // image_transformer_v2.py:52-54: 2x2 mean
      KDB_REQUIRE(h % 2 == 0 && w % 2 == 0, KDB_ERR_BAD_SHAPE, "token grid %dx%d at level %d is not even", h, w, l);
      std::vector<double> nxt((size_t)(h / 2) * (w / 2) * 2);
      for (int i = 0; i < h / 2; ++i)
        for (int j = 0; j < w / 2; ++j)
          for (int k = 0; k < 2; ++k)
            nxt[((size_t)i * (w / 2) + j) * 2 + k] =
                0.25 * (cur[((size_t)(2 * i) * w + 2 * j) * 2 + k] + cur[((size_t)(2 * i) * w + 2 * j + 1) * 2 + k] +
                        cur[((size_t)(2 * i + 1) * w + 2 * j) * 2 + k] + cur[((size_t)(2 * i + 1) * w + 2 * j + 1) * 2 + k]);
      cur.swap(nxt);
      h /= 2;
      w /= 2;
    }
  }
  // per-layer RoPE tables (freqs are per-layer buffers of the checkpoint)
  pt.rope.assign(m->n_layers, nullptr);
  {
    auto build = [&](const LayerPlan& L, int hl, int wl) -> int {
      if (L.attn_type == KDB_ATTN_NONE || L.e != 64) return 0;
      float2* tab = nullptr;
      int rc = dev_alloc(m, &tab, (size_t)hl * wl * L.nh * 16);
      if (rc) return rc;
      rc = launch_rope_table(pt.pos[L.level], L.freqs, tab, hl * wl, L.nh, L.e / 8, st);
      if (rc) return rc;
      pt.rope[L.exec_index] = tab;
      return 0;
    };
    const int nl = m->cfg.n_levels;
    for (int l = 0; l < nl; ++l) {
      const int hl = h0 >> l, wl = w0 >> l;
      int rc;
      if (l < nl - 1) {
        for (auto& L : m->down[l]) if ((rc = build(L, hl, wl))) return rc;
        for (auto& L : m->up[l]) if ((rc = build(L, hl, wl))) return rc;
      } else {
        for (auto& L : m->mid) if ((rc = build(L, hl, wl))) return rc;
      }
    }
    KDB_CUDA(cudaStreamSynchronize(st));
  }
  auto ins = m->pos_cache.emplace(key, std::move(pt));
  *out = &ins.first->second;
  return 0;
}

struct Workspace {
  std::vector<char*> xs, xup;
  char *xn = nullptr, *qkv = nullptr, *ao = nullptr, *hbuf = nullptr, *gbuf = nullptr, *mg = nullptr;
  float* rowss = nullptr;   // [tokens at level 0, SS_PARTS] sum(x^2) of the current residual stream (fused RMSNorm)
  size_t total = 0;
};

void carve(const KdbModelConfig& c, int prec, int B, int H, int W, char* base, Workspace& ws) {
  const size_t s = prec == KDB_PREC_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  const int n = c.n_levels;
  int64_t T = (int64_t)(H / c.patch_h) * (W / c.patch_w);
  size_t mx = 0, mqkv = 0, mh = 0, mg = 0;
  ws.xs.assign(n, nullptr);
  ws.xup.assign(n, nullptr);
  for (int l = 0; l < n; ++l) {
    const size_t xb = (size_t)B * T * c.width[l] * s;
    ws.xs[l] = take(xb);
    if (l < n - 1) ws.xup[l] = take(xb);
    mx = std::max(mx, xb);
    if (c.attn_type[l] != KDB_ATTN_NONE) mqkv = std::max(mqkv, 3 * xb);
    mh = std::max(mh, (size_t)B * T * 2 * c.d_ff[l] * s);
    if (l < n - 1) mg = std::max(mg, xb);
    T /= 4;
  }
  ws.xn = take(mx);
  ws.qkv = take(mqkv);
  ws.ao = take(mx);
  ws.hbuf = take(mh);
  ws.gbuf = take(mh / 2);
  ws.mg = take(mg);
  ws.rowss = reinterpret_cast<float*>(take((size_t)B * (H / c.patch_h) * (W / c.patch_w) * SS_PARTS * sizeof(float)));
  ws.total = off + 1024;
}

template <typename T>
int tap(KdbModel* m, const std::string& name, const T* p, int64_t n, cudaStream_t st) {
  if (m->tap_out == nullptr || m->tap_name != name) return 0;
  if (n > m->tap_cap) {
    m->tap_count = -n;
    return 0;
  }
  m->tap_count = n;
  return launch_to_f32<T>(p, m->tap_out, n, st);
}

template <typename T> struct WSel;
template <> struct WSel<float> {
  typedef float W;
  static const float* qkv(const LayerPlan& L) { return L.qkv_w; }
  static const float* out(const LayerPlan& L) { return L.out_w; }
  static const float* up(const LayerPlan& L) { return L.up_w; }
  static const float* down(const LayerPlan& L) { return L.down_w; }
  static const float* merge(const KdbModel* m, int l) { return m->merge_w[l]; }
  static const float* split(const KdbModel* m, int l) { return m->split_w[l]; }
};
template <> struct WSel<bf16> {
  typedef bf16 W;
  static const bf16* qkv(const LayerPlan& L) { return L.qkv_wb; }
  static const bf16* out(const LayerPlan& L) { return L.out_wb; }
  static const bf16* up(const LayerPlan& L) { return L.up_wb; }
  static const bf16* down(const LayerPlan& L) { return L.down_wb; }
  static const bf16* merge(const KdbModel* m, int l) { return m->merge_wb[l]; }
  static const bf16* split(const KdbModel* m, int l) { return m->split_wb[l]; }
};

// Linear dispatch: tensor-core kernel when the shape qualifies (bf16 only), SIMT otherwise.
template <typename T>
int linear(const T* A, const typename WSel<T>::W* W, T* C, int64_t M, int N, int K, const GemmEpi& epi, cudaStream_t st);
template <>
int linear<float>(const float* A, const float* W, float* C, int64_t M, int N, int K, const GemmEpi& epi, cudaStream_t st) {
  return launch_gemm_simt<float, float>(A, W, C, M, N, K, epi, st);
}
template <>
int linear<bf16>(const bf16* A, const bf16* W, bf16* C, int64_t M, int N, int K, const GemmEpi& epi, cudaStream_t st) {
  if (tc_gemm_supported(M, N, K, epi)) return launch_gemm_tc(A, W, C, M, N, K, epi, st);
  return launch_gemm_simt<bf16, bf16>(A, W, C, M, N, K, epi, st);
}

template <typename T>
int run_layer(KdbModel* m, const LayerPlan& L, T* x, int B, int h, int w, const PosTables* pt, const float* cond, int64_t cond_bs,
              Workspace& ws, cudaStream_t st) {
  const float* pos = pt->pos[L.level];
  const int64_t Ttok = (int64_t)h * w, M = (int64_t)B * Ttok;
  const int C = L.C;
  T* xn = reinterpret_cast<T*>(ws.xn);
  T* qkv = reinterpret_cast<T*>(ws.qkv);
  T* ao = reinterpret_cast<T*>(ws.ao);
  T* hb = reinterpret_cast<T*>(ws.hbuf);
  T* gb = reinterpret_cast<T*>(ws.gbuf);
  const std::string tag = "layer" + std::to_string(m->layer_counter++);
  int rc;
  // fused RMSNorm: possible when the whole batch shares one conditioning row (folded weights are per evaluation) and the GEMM
  // that produced x left its row statistics in ws.rowss
  const bool emit = std::is_same<T, bf16>::value && m->fuse_norm && C % 128 == 0;
  const bool fold = emit && cond_bs == 0 && m->fold_descs != nullptr;
  if (L.attn_type != KDB_ATTN_NONE) {
    GemmEpi qe;
    qe.mode = (L.e == 64 && pt->rope[L.exec_index] != nullptr) ? EPI_QKV_ROPE : EPI_STORE;
    qe.C = C;
    qe.nh = L.nh;
    qe.T_tokens = (int)Ttok;
    qe.rope = pt->rope[L.exec_index];
    qe.qk_scale = L.scale;
    GemmEpi qf = qe;
    qf.ss_in = ws.rowss;
    if (fold && m->ss_valid && L.qkv_wf != nullptr && tc_gemm_supported(M, 3 * C, C, qf)) {
      if ((rc = launch_gemm_tc(reinterpret_cast<const bf16*>(x), L.qkv_wf, reinterpret_cast<bf16*>(qkv), M, 3 * C, C, qf, st))) return rc;
      if (qf.mode == EPI_STORE && (rc = launch_qknorm_rope<T>(qkv, pos, L.freqs, L.scale, M, (int)Ttok, L.nh, L.e, st))) return rc;
    } else {
      if ((rc = launch_rmsnorm<T>(x, xn, cond + L.ada_attn, cond_bs, Ttok, M, C, st))) return rc;
      if ((rc = tap<T>(m, tag + ".xn1", xn, M * C, st))) return rc;
      if (std::is_same<T, bf16>::value && qe.mode == EPI_QKV_ROPE && tc_gemm_supported(M, 3 * C, C, qe)) {
        // cosine-sim scaling + RoPE fused into the qkv projection's epilogue
        if ((rc = launch_gemm_tc(reinterpret_cast<const bf16*>(xn), L.qkv_wb, reinterpret_cast<bf16*>(qkv), M, 3 * C, C, qe, st))) return rc;
      } else {
        if ((rc = linear<T>(xn, WSel<T>::qkv(L), qkv, M, 3 * C, C, GemmEpi{}, st))) return rc;
        if ((rc = launch_qknorm_rope<T>(qkv, pos, L.freqs, L.scale, M, (int)Ttok, L.nh, L.e, st))) return rc;
      }
    }
    if ((rc = tap<T>(m, tag + ".qkv", qkv, M * 3 * C, st))) return rc;
    // the bound holds for q, k normalised by the fused QKV epilogue or by qknorm_rope (both paths above)
    if ((rc = attention_dispatch<T>(qkv, ao, B, h, w, L.nh, L.e, L.attn_type, L.attn_param, L.shift, st, L.bounded ? L.scale : nullptr))) return rc;
    if ((rc = tap<T>(m, tag + ".ao", ao, M * C, st))) return rc;
    GemmEpi e;
    e.mode = EPI_RESID;
    e.resid = x;
    m->ss_valid = emit && tc_gemm_emits_rowss(M, C, C, e);
    if (m->ss_valid) e.ss_out = ws.rowss;
    if ((rc = linear<T>(ao, WSel<T>::out(L), x, M, C, C, e, st))) return rc;
    if ((rc = tap<T>(m, tag + ".attn", x, M * C, st))) return rc;
  }
  // fused feed-forward (128-wide levels): the hidden never leaves the SM.  Not when the hidden itself is being tapped.
  if (fold && m->ss_valid && L.up_wf != nullptr && L.down_wb != nullptr && tc_ffn_fused_supported(M, C, L.dff) &&
      !(m->tap_out != nullptr && m->tap_name == tag + ".geglu")) {
    if ((rc = launch_ffn_fused(reinterpret_cast<bf16*>(x), L.up_wf, L.down_wb, M, C, L.dff, ws.rowss, ws.rowss, st))) return rc;
    m->ss_valid = true;
    return tap<T>(m, tag + ".ff", x, M * C, st);
  }
  bool fused_geglu = false;
  if (fold && m->ss_valid && L.up_wf != nullptr && tc_gemm_geglu_supported(M, 2 * L.dff, C, true)) {
    if ((rc = launch_gemm_tc_geglu(reinterpret_cast<const bf16*>(x), L.up_wf, reinterpret_cast<bf16*>(gb), M, 2 * L.dff, C, st, ws.rowss))) return rc;
    fused_geglu = true;
  } else if ((rc = launch_rmsnorm<T>(x, xn, cond + L.ada_ff, cond_bs, Ttok, M, C, st))) {
    return rc;
  }
  if (!fused_geglu && std::is_same<T, bf16>::value && L.up_wb_il != nullptr && tc_gemm_geglu_supported(M, 2 * L.dff, C)) {
    if ((rc = launch_gemm_tc_geglu(reinterpret_cast<const bf16*>(xn), L.up_wb_il, reinterpret_cast<bf16*>(gb), M, 2 * L.dff, C, st)))
      return rc;
    fused_geglu = true;
  }
  if (!fused_geglu) {
    if ((rc = linear<T>(xn, WSel<T>::up(L), hb, M, 2 * L.dff, C, GemmEpi{}, st))) return rc;
    if ((rc = launch_geglu<T>(hb, gb, M, L.dff, st))) return rc;
  }
  if ((rc = tap<T>(m, tag + ".geglu", gb, M * L.dff, st))) return rc;
  GemmEpi e;
  e.mode = EPI_RESID;
  e.resid = x;
  m->ss_valid = emit && tc_gemm_emits_rowss(M, C, L.dff, e);
  if (m->ss_valid) e.ss_out = ws.rowss;
  if ((rc = linear<T>(gb, WSel<T>::down(L), x, M, C, L.dff, e, st))) return rc;
  return tap<T>(m, tag + ".ff", x, M * C, st);
}

template <typename T>
int forward_impl(KdbModel* m, int B, int H, int W, const float* x, const float* sigma, float sd, const float* cond, int64_t cond_bs,
                 float* out, Workspace& ws, cudaStream_t st) {
  const KdbModelConfig& c = m->cfg;
  const int n = c.n_levels;
  const int h0 = H / c.patch_h, w0 = W / c.patch_w;
  PosTables* pt = nullptr;
  int rc = ensure_pos(m, h0, w0, st, &pt);
  if (rc) return rc;
  m->layer_counter = 0;
  m->tap_count = 0;
  const float* patch_in_w = nullptr;
  const float *out_norm = nullptr, *patch_out_w = nullptr;
  GET("patch_in.proj.weight", &patch_in_w, c.width[0], (int64_t)c.patch_h * c.patch_w * c.in_channels);
  GET("out_norm.scale", &out_norm, c.width[0]);
  GET("patch_out.proj.weight", &patch_out_w, (int64_t)c.patch_h * c.patch_w * c.out_channels, c.width[0]);

  if (std::is_same<T, bf16>::value && m->fuse_norm && cond_bs == 0 && m->fold_descs != nullptr)
    if ((rc = launch_fold_norm_weights(m->fold_descs, m->n_fold, cond, st))) return rc;
  m->ss_valid = false;
  T* cur = reinterpret_cast<T*>(ws.xs[0]);
  if (std::is_same<T, bf16>::value && m->patch_in_wb != nullptr && tc_patch_in_supported(c.in_channels, c.patch_h, c.patch_w, c.width[0], W) &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    m->ss_valid = m->fuse_norm;
    if ((rc = launch_patch_in_tc(x, sigma, sd, m->patch_in_wb, reinterpret_cast<bf16*>(cur), B, H, W, c.width[0], m->ss_valid ? ws.rowss : nullptr, st)))
      return rc;
  } else if ((rc = launch_patch_in<T>(x, sigma, sd, patch_in_w, cur, B, c.in_channels, H, W, c.patch_h, c.patch_w, c.width[0], st))) {
    return rc;
  }
  if ((rc = tap<T>(m, "patch_in", cur, (int64_t)B * h0 * w0 * c.width[0], st))) return rc;

  int h = h0, w = w0;
  for (int l = 0; l < n - 1; ++l) {
    for (const LayerPlan& L : m->down[l])
      if ((rc = run_layer<T>(m, L, cur, B, h, w, pt, cond, cond_bs, ws, st))) return rc;
    if ((rc = tap<T>(m, "L" + std::to_string(l) + ".down", cur, (int64_t)B * h * w * c.width[l], st))) return rc;
    T* nxt = reinterpret_cast<T*>(ws.xs[l + 1]);
    GemmEpi me;                      // TokenMerge: the 2x2 gather rides on the GEMM's TMA loads when the geometry allows
    me.mC = c.width[l];
    me.mhc = h / 2;
    me.mwc = w / 2;
    const int64_t Mc = (int64_t)B * (h / 2) * (w / 2);
    bool merge_ss = false;
    if (std::is_same<T, bf16>::value && tc_gemm_supported(Mc, c.width[l + 1], 4 * c.width[l], me)) {
      merge_ss = m->fuse_norm && c.width[l + 1] % 128 == 0 && tc_gemm_emits_rowss(Mc, c.width[l + 1], 4 * c.width[l], me);
      if (merge_ss) me.ss_out = ws.rowss;      // row statistics of the merged tokens for the next level's first fused RMSNorm
      if ((rc = launch_gemm_tc(reinterpret_cast<const bf16*>(cur), m->merge_wb[l], reinterpret_cast<bf16*>(nxt), Mc, c.width[l + 1], 4 * c.width[l],
                               me, st)))
        return rc;
    } else {
      T* mg = reinterpret_cast<T*>(ws.mg);
      if ((rc = launch_merge_gather<T>(cur, mg, B, h, w, c.width[l], st))) return rc;
      if ((rc = linear<T>(mg, WSel<T>::merge(m, l), nxt, Mc, c.width[l + 1], 4 * c.width[l], GemmEpi{}, st))) return rc;
    }
    m->ss_valid = merge_ss;
    h /= 2;
    w /= 2;
    if ((rc = tap<T>(m, "L" + std::to_string(l) + ".merge", nxt, (int64_t)B * h * w * c.width[l + 1], st))) return rc;
    cur = nxt;
  }
  for (const LayerPlan& L : m->mid)
    if ((rc = run_layer<T>(m, L, cur, B, h, w, pt, cond, cond_bs, ws, st))) return rc;
  if ((rc = tap<T>(m, "mid", cur, (int64_t)B * h * w * c.width[n - 1], st))) return rc;
  for (int l = n - 2; l >= 0; --l) {
    T* up = reinterpret_cast<T*>(ws.xup[l]);
    GemmEpi e;
    e.mode = EPI_SPLIT_LERP;
    e.resid = ws.xs[l];
    e.fac = m->split_fac[l];
    e.hc = h;
    e.wc = w;
    e.C = c.width[l];
    m->ss_valid = std::is_same<T, bf16>::value && m->fuse_norm && tc_gemm_emits_rowss((int64_t)B * h * w, 4 * c.width[l], c.width[l + 1], e);
    if (m->ss_valid) e.ss_out = ws.rowss;
    if ((rc = linear<T>(cur, WSel<T>::split(m, l), up, (int64_t)B * h * w, 4 * c.width[l], c.width[l + 1], e, st))) return rc;
    h *= 2;
    w *= 2;
    if ((rc = tap<T>(m, "L" + std::to_string(l) + ".split", up, (int64_t)B * h * w * c.width[l], st))) return rc;
    for (const LayerPlan& L : m->up[l])
      if ((rc = run_layer<T>(m, L, up, B, h, w, pt, cond, cond_bs, ws, st))) return rc;
    if ((rc = tap<T>(m, "L" + std::to_string(l) + ".up", up, (int64_t)B * h * w * c.width[l], st))) return rc;
    cur = up;
  }
  if (std::is_same<T, bf16>::value && m->patch_out_wb != nullptr &&
      tc_patch_out_supported(c.width[0], c.out_channels, c.patch_h, c.patch_w, W)) {
    // out_norm as a row kernel, then the projection on the tensor core with un-patch + Karras combine in its epilogue
    if (m->ss_valid && m->patch_out_wf != nullptr && c.width[0] % 128 == 0 && c.width[0] <= 128 * SS_PARTS)
      return launch_patch_out_tc(reinterpret_cast<const bf16*>(cur), m->patch_out_wf, x, sigma, sd, out, B, H, W, c.width[0], st, ws.rowss);
    T* xn = reinterpret_cast<T*>(ws.xn);
    const int64_t M0 = (int64_t)B * h0 * w0;
    if ((rc = launch_rmsnorm<T>(cur, xn, out_norm, 0, M0, M0, c.width[0], st))) return rc;
    return launch_patch_out_tc(reinterpret_cast<const bf16*>(xn), m->patch_out_wb, x, sigma, sd, out, B, H, W, c.width[0], st);
  }
  return launch_patch_out<T>(cur, out_norm, patch_out_w, x, sigma, sd, out, B, c.out_channels, H, W, c.patch_h, c.patch_w, c.width[0], st);
}

}  // namespace

extern "C" {

int kdb_model_create(const KdbModelConfig* cfg, KdbModel** out) {
  KDB_REQUIRE(cfg && out, KDB_ERR_BAD_ARG, "model_create: NULL argument");
  KDB_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= KDB_MAX_LEVELS, KDB_ERR_BAD_ARG, "model_create: n_levels %d", cfg->n_levels);
  KDB_REQUIRE(cfg->patch_h >= 1 && cfg->patch_w >= 1 && cfg->in_channels >= 1 && cfg->out_channels >= 1, KDB_ERR_BAD_ARG,
              "model_create: bad patch/channels");
  KDB_REQUIRE(cfg->mapping_depth >= 0 && cfg->mapping_depth <= 8 && cfg->mapping_width >= 2, KDB_ERR_BAD_ARG, "model_create: bad mapping spec");
  for (int l = 0; l < cfg->n_levels; ++l) {
    KDB_REQUIRE(cfg->width[l] > 0 && cfg->depth[l] >= 0 && cfg->d_ff[l] > 0, KDB_ERR_BAD_ARG, "model_create: bad level %d", l);
    KDB_REQUIRE(cfg->attn_type[l] >= KDB_ATTN_NONE && cfg->attn_type[l] <= KDB_ATTN_SHIFTED_WINDOW, KDB_ERR_BAD_ARG,
                "model_create: unsupported self attention spec at level %d", l);
  }
  KdbModel* m = new KdbModel();
  m->cfg = *cfg;
  *out = m;
  return 0;
}

void kdb_model_destroy(KdbModel* m) {
  if (!m) return;
  free_owned(m);
  delete m;
}

int kdb_model_set_tensor(KdbModel* m, const char* key, const float* data, const int64_t* shape, int ndim) {
  KDB_REQUIRE(m && key && data && ndim >= 0 && ndim <= 4, KDB_ERR_BAD_ARG, "set_tensor: bad argument");
  TensorRef t;
  t.p = data;
  t.shape.assign(shape, shape + ndim);
  m->tensors[key] = t;
  m->finalized = false;
  return 0;
}

int kdb_model_finalize(KdbModel* m, void* stream) {
  KDB_REQUIRE(m, KDB_ERR_BAD_ARG, "finalize: NULL model");
  cudaStream_t st = (cudaStream_t)stream;
  free_owned(m);
  m->finalized = false;
  const KdbModelConfig& c = m->cfg;
  const int n = c.n_levels, mw = c.mapping_width;
  m->down.assign(n, {});
  m->up.assign(n, {});
  m->mid.clear();
  m->merge_w.assign(n, nullptr);
  m->split_w.assign(n, nullptr);
  m->split_fac.assign(n, nullptr);
  m->merge_wb.assign(n, nullptr);
  m->split_wb.assign(n, nullptr);
  int ada = 0, rc;
  // conditioning row order == execution order: down levels, mid, up levels (outermost last)
  for (int l = 0; l < n - 1; ++l) {
    m->down[l].resize(c.depth[l]);
    for (int i = 0; i < c.depth[l]; ++i) {
      rc = plan_layer(m, m->down[l][i], "down_levels." + std::to_string(l) + "." + std::to_string(i) + ".", l, i, &ada, st);
      if (rc) return rc;
    }
  }
  m->mid.resize(c.depth[n - 1]);
  for (int i = 0; i < c.depth[n - 1]; ++i) {
    rc = plan_layer(m, m->mid[i], "mid_level." + std::to_string(i) + ".", n - 1, i, &ada, st);
    if (rc) return rc;
  }
  for (int l = n - 2; l >= 0; --l) {
    m->up[l].resize(c.depth[l]);
    for (int i = 0; i < c.depth[l]; ++i) {   // image_transformer_v2.py:697: up-level layer index continues after the down level
      rc = plan_layer(m, m->up[l][i], "up_levels." + std::to_string(l) + "." + std::to_string(i) + ".", l, i + c.depth[l], &ada, st);
      if (rc) return rc;
    }
  }
  m->ada_total = ada;
  {
    int k = 0;
    for (int l = 0; l < n - 1; ++l)
      for (auto& L : m->down[l]) L.exec_index = k++;
    for (auto& L : m->mid) L.exec_index = k++;
    for (int l = n - 2; l >= 0; --l)
      for (auto& L : m->up[l]) L.exec_index = k++;
    m->n_layers = k;
    // table of (weights -> folded copy) pairs for the fused RMSNorm path
    std::vector<FoldDesc> descs;
    auto add = [&](const LayerPlan& L) {
      if (L.C % 8 != 0) return;
      if (L.qkv_wf != nullptr && L.ada_attn >= 0) descs.push_back(FoldDesc{L.qkv_wb, L.qkv_wf, 3 * L.C, L.C, L.ada_attn});
      if (L.up_wf != nullptr && L.up_wb_il != nullptr) descs.push_back(FoldDesc{L.up_wb_il, L.up_wf, 2 * L.dff, L.C, L.ada_ff});
    };
    for (int l = 0; l < n - 1; ++l)
      for (auto& L : m->down[l]) add(L);
    for (auto& L : m->mid) add(L);
    for (int l = n - 2; l >= 0; --l)
      for (auto& L : m->up[l]) add(L);
    m->n_fold = (int)descs.size();
    m->fold_descs = nullptr;
    if (!descs.empty()) {
      if ((rc = dev_alloc(m, &m->fold_descs, descs.size()))) return rc;
      KDB_CUDA(cudaMemcpyAsync(m->fold_descs, descs.data(), descs.size() * sizeof(FoldDesc), cudaMemcpyHostToDevice, st));
      KDB_CUDA(cudaStreamSynchronize(st));
    }
    const char* e = getenv("KDB200_NO_FUSED_NORM");
    m->fuse_norm = !(e != nullptr && e[0] == '1');
  }
  for (int l = 0; l < n - 1; ++l) {
    GET("merges." + std::to_string(l) + ".proj.weight", &m->merge_w[l], c.width[l + 1], 4 * c.width[l]);
    GET("splits." + std::to_string(l) + ".proj.weight", &m->split_w[l], 4 * c.width[l], c.width[l + 1]);
    GET("splits." + std::to_string(l) + ".fac", &m->split_fac[l], 1);
    if ((rc = make_bf16(m, m->merge_w[l], 4LL * c.width[l] * c.width[l + 1], &m->merge_wb[l], st))) return rc;
    if ((rc = make_bf16(m, m->split_w[l], 4LL * c.width[l] * c.width[l + 1], &m->split_wb[l], st))) return rc;
  }
  const float* tmp = nullptr;
  GET("patch_in.proj.weight", &tmp, c.width[0], (int64_t)c.patch_h * c.patch_w * c.in_channels);
  GET("out_norm.scale", &tmp, c.width[0]);
  GET("patch_out.proj.weight", &tmp, (int64_t)c.patch_h * c.patch_w * c.out_channels, c.width[0]);

  m->patch_in_wb = nullptr;
  if (c.in_channels == 3 && c.patch_h == 4 && c.patch_w == 4 && c.width[0] % 128 == 0) {
    const float* piw = nullptr;
    GET("patch_in.proj.weight", &piw, c.width[0], (int64_t)48);
    if ((rc = dev_alloc(m, &m->patch_in_wb, (size_t)c.width[0] * 64))) return rc;
    if ((rc = prepare_patch_in_weight(piw, m->patch_in_wb, c.width[0], st))) return rc;
  }
  {   // zero-padded bf16 patch_out weight [64, C0]
    const int Np = c.patch_h * c.patch_w * c.out_channels, C0 = c.width[0];
    m->patch_out_wb = nullptr;
    if (Np <= 64) {
      if ((rc = dev_alloc(m, &m->patch_out_wb, (size_t)64 * C0))) return rc;
      KDB_CUDA(cudaMemsetAsync(m->patch_out_wb, 0, (size_t)64 * C0 * sizeof(bf16), st));
      if ((rc = launch_f32_to_bf16(tmp, m->patch_out_wb, (int64_t)Np * C0, st))) return rc;
      m->patch_out_wf = nullptr;
      if (C0 % 8 == 0) {   // out_norm.scale is a plain parameter: fold it once
        const float* out_scale = nullptr;
        GET("out_norm.scale", &out_scale, C0);
        FoldDesc* d1 = nullptr;
        if ((rc = dev_alloc(m, &m->patch_out_wf, (size_t)64 * C0))) return rc;
        if ((rc = dev_alloc(m, &d1, 1))) return rc;
        const FoldDesc hd{m->patch_out_wb, m->patch_out_wf, 64, C0, 0};
        KDB_CUDA(cudaMemcpyAsync(d1, &hd, sizeof(FoldDesc), cudaMemcpyHostToDevice, st));
        if ((rc = launch_fold_norm_weights(d1, 1, out_scale, st))) return rc;
        KDB_CUDA(cudaStreamSynchronize(st));
      }
    }
  }
  // conditioning weights
  CondWeights& w = m->cw;
  w = CondWeights{};
  w.mw = mw;
  w.depth = c.mapping_depth;
  w.dff = c.mapping_d_ff;
  w.n_classes = c.num_classes;
  w.mcond_dim = c.mapping_cond_dim;
  w.ada_total = ada;
  GET("time_emb.weight", &w.time_emb, mw / 2, 1);
  GET("time_in_proj.weight", &w.time_in, mw, mw);
  GET("aug_emb.weight", &w.aug_emb, mw / 2, 9);
  GET("aug_in_proj.weight", &w.aug_in, mw, mw);
  if (c.num_classes > 0) GET("class_emb.weight", &w.class_emb, c.num_classes, mw);
  if (c.mapping_cond_dim > 0) GET("mapping_cond_in_proj.weight", &w.mcond_in, mw, c.mapping_cond_dim);
  GET("mapping.in_norm.scale", &w.in_norm, mw);
  GET("mapping.out_norm.scale", &w.out_norm, mw);
  for (int i = 0; i < c.mapping_depth; ++i) {
    const std::string p = "mapping.blocks." + std::to_string(i) + ".";
    GET(p + "norm.scale", &w.blk_norm[i], mw);
    GET(p + "up_proj.weight", &w.blk_up[i], 2 * c.mapping_d_ff, mw);
    GET(p + "down_proj.weight", &w.blk_down[i], mw, c.mapping_d_ff);
  }
  // concatenated AdaRMSNorm projection [ada_total, mw]
  if ((rc = dev_alloc(m, &m->ada_cat, (size_t)ada * mw))) return rc;
  auto put = [&](const LayerPlan& L) -> int {
    if (L.ada_attn >= 0)
      KDB_CUDA(cudaMemcpyAsync(m->ada_cat + (size_t)L.ada_attn * mw, L.attn_norm_w, sizeof(float) * L.C * mw, cudaMemcpyDeviceToDevice, st));
    KDB_CUDA(cudaMemcpyAsync(m->ada_cat + (size_t)L.ada_ff * mw, L.ff_norm_w, sizeof(float) * L.C * mw, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  for (auto& lv : m->down)
    for (auto& L : lv)
      if ((rc = put(L))) return rc;
  for (auto& L : m->mid)
    if ((rc = put(L))) return rc;
  for (auto& lv : m->up)
    for (auto& L : lv)
      if ((rc = put(L))) return rc;
  w.ada_cat = m->ada_cat;
  KDB_CUDA(cudaStreamSynchronize(st));
  m->finalized = true;
  return 0;
}

int64_t kdb_model_cond_stride(const KdbModel* m) {
  if (!m || !m->finalized) return KDB_ERR_NOT_FINAL;
  return (int64_t)align_up((size_t)(m->ada_total + m->cfg.mapping_width), 4);
}

int kdb_model_conditioning(KdbModel* m, int rows, const float* sigma, const float* aug_cond, const int64_t* class_cond,
                           const float* mapping_cond, float* cond_out, void* stream) {
  KDB_REQUIRE(m && m->finalized, KDB_ERR_NOT_FINAL, "conditioning: model not finalized");
  KDB_REQUIRE(rows > 0 && sigma && cond_out, KDB_ERR_BAD_ARG, "conditioning: bad arguments");
  // image_transformer_v2.py:729-732
  KDB_REQUIRE(!(m->cfg.num_classes > 0 && class_cond == nullptr), KDB_ERR_BAD_ARG, "class_cond must be specified if num_classes > 0");
  KDB_REQUIRE(!(m->cfg.mapping_cond_dim > 0 && mapping_cond == nullptr), KDB_ERR_BAD_ARG,
              "mapping_cond must be specified if mapping_cond_dim > 0");
  return launch_conditioning(m->cw, rows, sigma, aug_cond, class_cond, mapping_cond, cond_out, kdb_model_cond_stride(m),
                             (cudaStream_t)stream);
}

size_t kdb_model_workspace_bytes(const KdbModel* m, int precision, int batch, int height, int width) {
  if (!m || batch <= 0 || height <= 0 || width <= 0) return 0;
  Workspace ws;
  carve(m->cfg, precision, batch, height, width, nullptr, ws);
  return ws.total;
}

int kdb_model_forward(KdbModel* m, int precision, int batch, int height, int width, const float* x, const float* sigma,
                      float sigma_data, const float* cond, int64_t cond_batch_stride, float* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  KDB_REQUIRE(m && m->finalized, KDB_ERR_NOT_FINAL, "forward: model not finalized");
  KDB_REQUIRE(x && sigma && cond && out && workspace && batch > 0, KDB_ERR_BAD_ARG, "forward: NULL argument");
  KDB_REQUIRE(precision == KDB_PREC_FP32 || precision == KDB_PREC_BF16, KDB_ERR_BAD_ARG, "forward: bad precision %d", precision);
  const KdbModelConfig& c = m->cfg;
  KDB_REQUIRE(height % c.patch_h == 0 && width % c.patch_w == 0, KDB_ERR_BAD_SHAPE, "forward: %dx%d not divisible by the patch size", height, width);
  const int div = 1 << (c.n_levels - 1);
  KDB_REQUIRE((height / c.patch_h) % div == 0 && (width / c.patch_w) % div == 0, KDB_ERR_BAD_SHAPE,
              "forward: token grid %dx%d not divisible by 2^(levels-1)", height / c.patch_h, width / c.patch_w);
  KDB_REQUIRE(!(sigma_data > 0.f && c.in_channels != c.out_channels), KDB_ERR_BAD_ARG, "forward: preconditioning needs C_in == C_out");
  Workspace ws;
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  carve(c, precision, batch, height, width, base, ws);
  KDB_REQUIRE(ws.total <= workspace_bytes, KDB_ERR_WORKSPACE, "forward: workspace %zu < required %zu", workspace_bytes, ws.total);
  int rc;
  if (precision == KDB_PREC_FP32)
    rc = forward_impl<float>(m, batch, height, width, x, sigma, sigma_data, cond, cond_batch_stride, out, ws, (cudaStream_t)stream);
  else
    rc = forward_impl<bf16>(m, batch, height, width, x, sigma, sigma_data, cond, cond_batch_stride, out, ws, (cudaStream_t)stream);
  m->tap_out = nullptr;
  m->tap_name.clear();
  return rc;
}

int kdb_model_debug_tap(KdbModel* m, const char* name, float* out, int64_t capacity) {
  KDB_REQUIRE(m && name && out && capacity > 0, KDB_ERR_BAD_ARG, "debug_tap: bad argument");
  m->tap_name = name;
  m->tap_out = out;
  m->tap_cap = capacity;
  m->tap_count = 0;
  return 0;
}

int64_t kdb_model_tap_count(const KdbModel* m) { return m ? m->tap_count : 0; }

int kdb_attention(int precision, int fast, const void* qkv, void* out, int batch, int h, int w, int n_heads, int d_head, int attn_type,
                  int attn_param, int shift, const float* logit_bound, void* stream) {
  KDB_REQUIRE(qkv && out && batch > 0 && h > 0 && w > 0 && n_heads > 0 && d_head > 0, KDB_ERR_BAD_ARG, "attention: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == KDB_PREC_FP32) {
    KDB_REQUIRE(!fast, KDB_ERR_UNSUPPORTED, "attention: tensor-core path is bf16 only");
    return launch_attention_generic<float>(static_cast<const float*>(qkv), static_cast<float*>(out), batch, h, w, n_heads, d_head,
                                           attn_type, attn_param, shift, st);
  }
  if (fast) {
    KDB_REQUIRE(tc_attention_supported(h, w, n_heads, d_head, attn_type, attn_param), KDB_ERR_UNSUPPORTED,
                "attention: shape not covered by the tensor-core kernels");
    return launch_attention_tc(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), batch, h, w, n_heads, d_head, attn_type,
                               attn_param, shift, st, logit_bound);
  }
  return launch_attention_generic<bf16>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), batch, h, w, n_heads, d_head, attn_type,
                                        attn_param, shift, st);
}

}  // extern "C"
