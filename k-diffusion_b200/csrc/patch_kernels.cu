// patch_kernels.cu -- tiled patch-in / patch-out kernels (64 tokens per CTA, coalesced pixel and token I/O).
//   patch_in : x[B,C,H,W] * c_in(sigma) -> 'b (h nh) (w nw) c -> b h w (nh nw c)' -> Linear(K=ph*pw*C -> N)
//              (reference image_transformer_v2.py:586-595,723-724 and layers.py:88-90 for c_in)
//   patch_out: RMSNorm(out_norm) -> Linear(C0 -> ph*pw*Cout) -> 'b h w (nh nw c) -> b c (h nh) (w nw)' -> c_out*F + c_skip*x
//              (reference :598-607,758-760 and layers.py:88-90)
// Both are <0.5 % of the model's MACs: fp32 FMA kernels whose job is to stay HBM-bound.
#include <algorithm>

#include "model_kernels.cuh"

namespace kdb {
namespace {

constexpr int TOK = 64;
constexpr float kEps = 1e-6f;

struct PatchGeom {
  int C, H, W, ph, pw, th, tw;   // image channels / size, patch, token grid
  int64_t tokens;
};

__device__ __forceinline__ void tok_coords(const PatchGeom& g, int64_t tok, int& b, int& ty, int& tx) {
  const int64_t per = (int64_t)g.th * g.tw;
  b = (int)(tok / per);
  const int r = (int)(tok - (int64_t)b * per);
  ty = r / g.tw;
  tx = r - ty * g.tw;
}
// coordinates of token (tile origin + t) from the tile origin's coordinates, without 64-bit division
__device__ __forceinline__ void tok_step(const PatchGeom& g, int b0, int ty0, int tx0, int t, int& b, int& ty, int& tx) {
  b = b0; ty = ty0; tx = tx0 + t;
  while (tx >= g.tw) { tx -= g.tw; ++ty; }
  while (ty >= g.th) { ty -= g.th; ++b; }
}

template <typename T> __device__ __forceinline__ void store_pair(T* o, float a, float b);
template <> __device__ __forceinline__ void store_pair<float>(float* o, float a, float b) { *reinterpret_cast<float2*>(o) = make_float2(a, b); }
template <> __device__ __forceinline__ void store_pair<bf16>(bf16* o, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(a, b); }

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patch_in_tiled(const float* __restrict__ x, const float* __restrict__ sigma, float sd,
                                                      const float* __restrict__ W, T* __restrict__ out, PatchGeom g, int N) {
  extern __shared__ float sm[];
  const int K = g.ph * g.pw * g.C;
  float* patch = sm;                 // [K][TOK]
  float* Ws = sm + K * TOK;          // [K][N]
  for (int idx = threadIdx.x; idx < K * N; idx += 256) {   // W [N][K] -> Ws [K][N], once per CTA
    const int n = idx / K, k = idx - n * K;
    Ws[k * N + n] = __ldg(W + idx);
  }
  const int pairs = N / 2;
  const int groups = 256 / pairs;
  const int per_group = TOK / groups;     // <= 32
  const int p = threadIdx.x % pairs, tg = threadIdx.x / pairs;
  const int64_t n_tiles = (g.tokens + TOK - 1) / TOK;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tok0 = tile * TOK;
    int b0, ty0, tx0;
    tok_coords(g, tok0, b0, ty0, tx0);
    __syncthreads();                      // previous tile's patch fully consumed (and Ws visible on the first pass)
    // gather pixels, pixel-contiguous order: idx = ((c*ph + nh)*TOK + t)*pw + nw
    for (int idx = threadIdx.x; idx < K * TOK; idx += 256) {
      const int nw = idx % g.pw;
      const int t = (idx / g.pw) % TOK;
      const int rest = idx / (g.pw * TOK);
      const int nh = rest % g.ph, c = rest / g.ph;
      float v = 0.f;
      if (tok0 + t < g.tokens) {
        int b, ty, tx;
        tok_step(g, b0, ty0, tx0, t, b, ty, tx);
        float c_in = 1.f;
        if (sd > 0.f) {
          float cs, co;
          karras_scalings(__ldg(sigma + b), sd, cs, co, c_in);
        }
        v = __ldg(x + (((int64_t)b * g.C + c) * g.H + (ty * g.ph + nh)) * g.W + (tx * g.pw + nw)) * c_in;
      }
      patch[((nh * g.pw + nw) * g.C + c) * TOK + t] = v;
    }
    __syncthreads();
    float a0[32], a1[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a0[i] = a1[i] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float2 w = *reinterpret_cast<const float2*>(Ws + k * N + 2 * p);
      const float* pr = patch + k * TOK + tg * per_group;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < per_group) {
          const float a = pr[i];
          a0[i] = fmaf(a, w.x, a0[i]);
          a1[i] = fmaf(a, w.y, a1[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per_group) {
        const int64_t tok = tok0 + tg * per_group + i;
        if (tok < g.tokens) store_pair<T>(out + tok * N + 2 * p, a0[i], a1[i]);
      }
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patch_out_tiled(const T* __restrict__ tokens, const float* __restrict__ nscale,
                                                       const float* __restrict__ W, const float* __restrict__ x_in,
                                                       const float* __restrict__ sigma, float sd, float* __restrict__ out, PatchGeom g,
                                                       int C0) {
  extern __shared__ float sm[];
  const int N = g.ph * g.pw * g.C;                  // g.C = output channels here
  float* xn = sm;                                   // [C0][TOK+1]
  float* Ws = xn + C0 * (TOK + 1);                  // [C0][N]
  float* ys = Ws + C0 * N;                          // [N][TOK+1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < N * C0; idx += 256) {      // W [N][C0] -> Ws [C0][N], once per CTA
    const int n = idx / C0, k = idx - n * C0;
    Ws[k * N + n] = __ldg(W + idx);
  }
  const int64_t n_tiles = (g.tokens + TOK - 1) / TOK;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tok0 = tile * TOK;
    int b0, ty0, tx0;
    tok_coords(g, tok0, b0, ty0, tx0);
    __syncthreads();                       // previous tile's xn / ys consumed
    // RMSNorm: warp per token (8 tokens per warp), normalised row written transposed
    for (int t = warp; t < TOK; t += 8) {
      const int64_t tok = tok0 + t;
      float ss = 0.f;
      if (tok < g.tokens)
        for (int c = lane; c < C0; c += 32) {
          const float v = to_f(tokens[tok * C0 + c]);
          ss = fmaf(v, v, ss);
        }
      ss = warp_sum(ss);
      const float rstd = rsqrtf(ss / (float)C0 + kEps);
      for (int c = lane; c < C0; c += 32) {
        const float v = tok < g.tokens ? to_f(tokens[tok * C0 + c]) : 0.f;
        xn[c * (TOK + 1) + t] = to_f(from_f<T>(v * (__ldg(nscale + c) * rstd)));
      }
    }
    __syncthreads();
    // thread = (token, group of output features)
    const int t = threadIdx.x % TOK, gq = threadIdx.x / TOK;      // 4 groups
    const int per = (N + 3) / 4;
    for (int nb = gq * per; nb < min(N, (gq + 1) * per); nb += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      const int cnt = min(16, min(N, (gq + 1) * per) - nb);
      for (int k = 0; k < C0; ++k) {
        const float a = xn[k * (TOK + 1) + t];
        const float* wr = Ws + k * N + nb;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < cnt) acc[j] = fmaf(a, wr[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < cnt) ys[(nb + j) * (TOK + 1) + t] = to_f(from_f<T>(acc[j]));
    }
    __syncthreads();
    // scatter in pixel-contiguous order: idx = ((c*ph + nh)*TOK + t)*pw + nw
    for (int idx = threadIdx.x; idx < N * TOK; idx += 256) {
      const int nw = idx % g.pw;
      const int tt = (idx / g.pw) % TOK;
      const int rest = idx / (g.pw * TOK);
      const int nh = rest % g.ph, c = rest / g.ph;
      if (tok0 + tt >= g.tokens) continue;
      int b, ty, tx;
      tok_step(g, b0, ty0, tx0, tt, b, ty, tx);
      const float y = ys[((nh * g.pw + nw) * g.C + c) * (TOK + 1) + tt];
      const int64_t o = (((int64_t)b * g.C + c) * g.H + (ty * g.ph + nh)) * g.W + (tx * g.pw + nw);
      if (sd > 0.f) {
        float c_skip, c_out, c_in;
        karras_scalings(__ldg(sigma + b), sd, c_skip, c_out, c_in);
        out[o] = y * c_out + __ldg(x_in + o) * c_skip;
      } else {
        out[o] = y;
      }
    }
  }
}

}  // namespace

template <typename T>
bool launch_patch_in_tiled(const float* x, const float* sigma, float sigma_data, const float* W, T* out, int B, int C, int H, int Wd, int ph,
                           int pw, int N, cudaStream_t st, int* rc) {
  const int K = ph * pw * C;
  const int pairs = N / 2;
  const size_t smem = sizeof(float) * ((size_t)K * TOK + (size_t)K * N);
  if (N % 2 != 0 || pairs > 256 || 256 % pairs != 0 || TOK / (256 / pairs) > 32 || smem > 160 * 1024) return false;
  PatchGeom g{C, H, Wd, ph, pw, H / ph, Wd / pw, (int64_t)B * (H / ph) * (Wd / pw)};
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(patch_in_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  patch_in_tiled<T><<<(unsigned)std::min<int64_t>(ceil_div(g.tokens, TOK), kNumSMs * 2), 256, smem, st>>>(x, sigma, sigma_data, W, out, g, N);
  count_launch(F_PATCH_IN, st);
  cudaError_t e = cudaGetLastError();
  *rc = e == cudaSuccess ? 0 : cuda_fail(e, "patch_in_tiled");
  return true;
}
template bool launch_patch_in_tiled<float>(const float*, const float*, float, const float*, float*, int, int, int, int, int, int, int, cudaStream_t, int*);
template bool launch_patch_in_tiled<bf16>(const float*, const float*, float, const float*, bf16*, int, int, int, int, int, int, int, cudaStream_t, int*);

template <typename T>
bool launch_patch_out_tiled(const T* tokens, const float* norm_scale, const float* W, const float* x_in, const float* sigma, float sigma_data,
                            float* out, int B, int Cout, int H, int Wd, int ph, int pw, int C0, cudaStream_t st, int* rc) {
  const int N = ph * pw * Cout;
  const size_t smem = sizeof(float) * ((size_t)C0 * (TOK + 1) + (size_t)C0 * N + (size_t)N * (TOK + 1));
  if (N > 64 || smem > 160 * 1024) return false;
  PatchGeom g{Cout, H, Wd, ph, pw, H / ph, Wd / pw, (int64_t)B * (H / ph) * (Wd / pw)};
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(patch_out_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  patch_out_tiled<T><<<(unsigned)std::min<int64_t>(ceil_div(g.tokens, TOK), kNumSMs * 2), 256, smem, st>>>(tokens, norm_scale, W, x_in, sigma, sigma_data, out, g, C0);
  count_launch(F_PATCH_OUT, st);
  cudaError_t e = cudaGetLastError();
  *rc = e == cudaSuccess ? 0 : cuda_fail(e, "patch_out_tiled");
  return true;
}
template bool launch_patch_out_tiled<float>(const float*, const float*, const float*, const float*, const float*, float, float*, int, int, int,
                                            int, int, int, int, cudaStream_t, int*);
template bool launch_patch_out_tiled<bf16>(const bf16*, const float*, const float*, const float*, const float*, float, float*, int, int, int,
                                           int, int, int, int, cudaStream_t, int*);

}  // namespace kdb
