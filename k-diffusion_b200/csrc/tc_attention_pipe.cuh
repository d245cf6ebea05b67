// tc_attention_pipe.cuh -- persistent, pipelined attention for cosine-similarity logits: global, 7x7 neighbourhood and shifted-window
// attention on one skeleton (included inside tc_attention.cu's anonymous namespace).  Reference: image_transformer_v2.py:355-396
// (global), :399-443 (neighbourhood via natten), :253-337,446-476 (shifted window); d_head 64.
//
// Why.  attn_tc_kernel is a one-shot CTA: load -> S = Q K^T -> softmax -> O = P V -> store, nothing in flight behind it, and for
// several key blocks it walks the keys twice (row maxima, then exp / P V).  Measured (round 1 / round 2 call 1): global S = 256
// 26.7 us against a 5 us HBM floor, window level 0 41-48 us against 20 us, neighbourhood level 0 175 us at batch 32.  Two facts
// remove all of that:
//
//  1. BOUNDED LOGITS.  q and k arrive cosine-normalised: |q| = |k| = sqrt(scale_h) (reference :106-114; RoPE is a rotation), so
//     |q . k| <= scale_h.  softmax is shift invariant, so exp(s - scale_h) with the FIXED shift scale_h is exact softmax
//     arithmetic -- no running maximum, no accumulator rescale, no first pass; masked keys simply get p = 0.  The smallest term is
//     exp(-2 scale_h): the host selects this kernel only for scale_h <= 40 (fp32 / bf16 exponent range).
//  2. d_head = 64 makes attention MUFU-bound, not tensor-bound: a 128 x 128 logit tile is 512 cycles of tcgen05.mma (4 x 64 for
//     Q K^T, 8 x 32 for P V) but 16384 exponentials = 1024 cycles of the SM's 16-lane MUFU unit.  So the kernel is built to keep
//     the exponential pipe (and, for the windowed modes, the TMA queue) busy all the time and to hide everything else behind it.
//
// One CTA per SM, 320 threads, loops over PAIRS of 128-row query tiles:
//   warps 0-3   softmax group 0: tile 0 of the pair, thread = query row (TMEM lane), S0 / O0
//   warps 4-7   softmax group 1: tile 1 of the pair, S1 / O1
//   warp  8     TMA producer: the pair's Q tiles (double buffered across pairs) and a 3-stage K/V ring that runs ahead across pairs
//   warp  9     tcgen05.mma issuer
// Per key block j the issuer runs, for t = 0, 1:  wait P_t(j) -> O_t += P_t V -> S_t = Q_t K(j+1)^T, so while group t exponentiates
// block j the tensor core finishes the other group's block and the next S tile is ready the moment a group asks for it.
//   GLOBAL : pair = 256 queries of one (image, head); both tiles share each K/V stage; nb = S / 128 key blocks
//   NA     : tile = 8 x 16 query block of one head; its own K/V = clamped 14 x 22 halo in 3 blocks of 5 halo rows (110 keys);
//            the 7 x 7 neighbourhood is a 128-bit mask per (row, block); 32-key chunks no row of the warp needs are skipped
//   WINDOW : tile = one 8 x 8 window x 2 heads (rows = head-major); one key block; S is 128 x 128 with the two 64 x 64 diagonal
//            blocks in use; the roll is TMA coordinates, the seam mask two 16-key groups per chunk
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384).
// P (the bf16 probabilities, A operand of the P V MMA) has two homes, template flag PT:
//   PT = true  (default): P stays in TENSOR MEMORY.  The softmax thread packs two probabilities per 32-bit column and writes them
//                with tcgen05.st over the first 64 columns of its own S tile (chunk c of P lands on columns the thread has already
//                read), and the P V MMA takes its A operand from TMEM.  No shared-memory round trip, no proxy fence, and the
//                64 KiB that P would need buy two more K/V stages: 5 x 32 KiB in flight instead of 3 -- the windowed modes consume
//                one stage per tile and were bound by load latency with 3.
//   PT = false (KDB200_ATTN_P_SMEM=1): P in shared memory (2 x 32 KiB, UMMA K-major SW128 layout), 3 K/V stages.
// Shared memory: Q 4 x 16 KiB + K/V stages x 32 KiB (+ P 2 x 32 KiB) = 224 KiB either way.
#pragma once

constexpr int PA_MAX_STAGES = 5;
constexpr size_t PA_SMEM = (size_t)14 * TILE_BYTES + 1024 + 512;

struct PipeAttnBars {
  uint64_t q_full[2], q_empty[2];
  uint64_t kv_full[PA_MAX_STAGES], kv_empty[PA_MAX_STAGES];
  uint64_t s_ready[2], p_ready[2], pv_done[2];
  uint32_t tmem;
};

struct PipeAttnParams {
  bf16* out;
  const float* bound;      // [nh] upper bound of |q . k| per head (= the layer's cosine-similarity scale)
  int B, h, w, nh, shift;
  int nb;                  // key blocks per tile
  int n_pairs;             // tile pairs in total
  long long* trace;        // debug (KDB200_ATTN_TRACE=1): [4 roles][PA_TRACE_N] (code, clock64) of CTA 0, else nullptr
};
constexpr int PA_TRACE_N = 192;
// roles: 0 producer, 1 MMA issuer, 2 / 3 softmax group 0 / 1 (row 0 of the group)
#define PA_TRACE(role_, code_)                                                                                     \
  do {                                                                                                             \
    if (p.trace != nullptr && blockIdx.x == 0 && tr_n < PA_TRACE_N) {                                              \
      p.trace[((role_) * PA_TRACE_N + tr_n) * 2] = (code_);                                                        \
      p.trace[((role_) * PA_TRACE_N + tr_n) * 2 + 1] = clock64();                                                  \
      ++tr_n;                                                                                                      \
    }                                                                                                              \
  } while (0)

struct PipeTile {          // decoded query tile
  int b, head;             // WINDOW: head = first head of the pair
  int q0;                  // GLOBAL: first query token
  int qi0, qj0, r0, c0;    // NA: query block origin, clamped halo origin.  WINDOW: qi0 = wi, qj0 = wj
};

template <int MODE>
__device__ __forceinline__ PipeTile pipe_decode(const PipeAttnParams& p, int pair, int t) {
  PipeTile x{};
  if constexpr (MODE == MODE_GLOBAL) {
    const int qpairs = (p.h * p.w) >> 8;
    x.q0 = (pair % qpairs) * 256 + t * 128;
    pair /= qpairs;
    x.head = pair % p.nh;
    x.b = pair / p.nh;
  } else if constexpr (MODE == MODE_NA) {
    const int nbw = p.w / NA_QW, nblk = (p.h / NA_QH) * nbw;
    int g = pair * 2 + t;
    const int blk = g % nblk;
    g /= nblk;
    x.head = g % p.nh;
    x.b = g / p.nh;
    x.qi0 = (blk / nbw) * NA_QH;
    x.qj0 = (blk % nbw) * NA_QW;
    x.r0 = min(max(x.qi0 - 3, 0), p.h - NA_KH);
    x.c0 = min(max(x.qj0 - 3, 0), p.w - NA_KW);
  } else {
    const int nww = p.w / 8, nwh = p.h / 8, hp = p.nh / 2;
    int g = pair * 2 + t;
    x.qj0 = g % nww;
    g /= nww;
    x.qi0 = g % nwh;
    g /= nwh;
    x.head = (g % hp) * 2;
    x.b = g / hp;
  }
  return x;
}

template <int MODE, bool PT>
__global__ void __launch_bounds__(320, 1) attn_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_kv,
                                                           const PipeAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool SHARED_KV = MODE == MODE_GLOBAL;
  constexpr int PA_STAGES = PT ? 5 : 3;
  int tr_n = 0;
  constexpr uint32_t KV_BYTES = (MODE == MODE_NA) ? NA_BLK_KEYS * 128 : TILE_BYTES;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = base;                                   // [qbuf 2][tile 2] x 16 KiB
  uint8_t* sKV = sQ + 4 * TILE_BYTES;                   // [stage][K | V] x 16 KiB
  uint8_t* sP = sKV + 2 * PA_STAGES * TILE_BYTES;       // [tile 2][key half 2] x 16 KiB
  PipeAttnBars* bars = reinterpret_cast<PipeAttnBars*>(sP + (PT ? 0 : 4) * TILE_BYTES);      // 14 tiles either way
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = p.nh, nb = p.nb;
  const int n_local = (int)blockIdx.x < p.n_pairs ? (p.n_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap);
    if constexpr (MODE == MODE_NA) tc::tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars->q_full[i], 1);
      tc::mbar_init(&bars->q_empty[i], 1);
      tc::mbar_init(&bars->s_ready[i], 1);
      tc::mbar_init(&bars->p_ready[i], 128);
      tc::mbar_init(&bars->pv_done[i], 1);
    }
    for (int i = 0; i < PA_STAGES; ++i) {
      tc::mbar_init(&bars->kv_full[i], 1);
      tc::mbar_init(&bars->kv_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  if constexpr (MODE == MODE_NA) {
    // rows 110..127 of every V tile are never written by TMA: they must be finite (P there is exactly 0)
    for (int i = threadIdx.x; i < PA_STAGES * (ROWS - NA_BLK_KEYS) * 8; i += blockDim.x) {
      const int st = i / ((ROWS - NA_BLK_KEYS) * 8), r = i - st * ((ROWS - NA_BLK_KEYS) * 8);
      *reinterpret_cast<uint4*>(sKV + (size_t)(st * 2 + 1) * TILE_BYTES + NA_BLK_KEYS * 128 + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async();
  }
  if constexpr (MODE == MODE_WINDOW && !PT) {
    // P of a window tile is block diagonal (rows of head 0 x keys of head 0, head 1 x head 1): the off-diagonal halves are zero for
    // every tile of this kernel -- written once here, never touched again
    if (warp < 8) {
      const int t = warp >> 2, row = (warp & 3) * 32 + lane, hd = row >> 6;
      uint8_t* other = sP + (size_t)(t * 2 + (1 - hd)) * TILE_BYTES;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint4*>(other + p_offset(row, jj)) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async();
  }
  if (warp == 9) tc::tmem_alloc(&bars->tmem, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  tc::pdl_wait();                    // the qkv projection before us must be complete from here on
  tc::pdl_launch_dependents();
  const uint32_t tmem = bars->tmem;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      uint32_t st = 0, ph = 0;
      auto load_window_tile = [&](uint8_t* dst, int which, const PipeTile& x, uint64_t* bar) {
        if (p.shift == 0) {   // unshifted window: one 8x8 box per head, rows in (lr, lc) order
#pragma unroll
          for (int hd = 0; hd < 2; ++hd) tc::tma_load_4d(dst + hd * 64 * 128, &tmap, bar, (which * nh + x.head + hd) * DH, x.qj0 * 8, x.qi0 * 8, x.b);
          return;
        }
#pragma unroll
        for (int hd = 0; hd < 2; ++hd)
#pragma unroll
          for (int quad = 0; quad < 4; ++quad) {
            const int rr = (x.qi0 * 8 + (quad >> 1) * 4 - p.shift + p.h) % p.h;   // rolled -> original coordinates (:274)
            const int cc = (x.qj0 * 8 + (quad & 1) * 4 - p.shift + p.w) % p.w;
            tc::tma_load_4d(dst + (hd * 64 + quad * 16) * 128, &tmap, bar, (which * nh + x.head + hd) * DH, cc, rr, x.b);
          }
      };
      auto kv_stage = [&](const PipeTile& x, int j) {
        tc::mbar_wait(&bars->kv_empty[st], ph ^ 1u);
        tc::mbar_arrive_expect_tx(&bars->kv_full[st], 2 * KV_BYTES);
        PA_TRACE(0, 100 + j);
        uint8_t* k = sKV + (size_t)(st * 2) * TILE_BYTES;
        if constexpr (MODE == MODE_GLOBAL) {
          tc::tma_load_3d(k, &tmap, &bars->kv_full[st], (nh + x.head) * DH, j * ROWS, x.b);
          tc::tma_load_3d(k + TILE_BYTES, &tmap, &bars->kv_full[st], (2 * nh + x.head) * DH, j * ROWS, x.b);
        } else if constexpr (MODE == MODE_NA) {
          tc::tma_load_4d(k, &tmap_kv, &bars->kv_full[st], (nh + x.head) * DH, x.c0, x.r0 + j * NA_BLK_ROWS, x.b);
          tc::tma_load_4d(k + TILE_BYTES, &tmap_kv, &bars->kv_full[st], (2 * nh + x.head) * DH, x.c0, x.r0 + j * NA_BLK_ROWS, x.b);
        } else {
          load_window_tile(k, 1, x, &bars->kv_full[st]);
          load_window_tile(k + TILE_BYTES, 2, x, &bars->kv_full[st]);
        }
        if (++st == PA_STAGES) {
          st = 0;
          ph ^= 1u;
        }
      };
      for (int u = 0; u < n_local; ++u) {
        const int pair = (int)blockIdx.x + u * (int)gridDim.x;
        const PipeTile x0 = pipe_decode<MODE>(p, pair, 0), x1 = pipe_decode<MODE>(p, pair, 1);
        const int qb = u & 1;
        tc::mbar_wait(&bars->q_empty[qb], (uint32_t)(((u >> 1) & 1) ^ 1));
        tc::mbar_arrive_expect_tx(&bars->q_full[qb], 2 * TILE_BYTES);
        PA_TRACE(0, 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const PipeTile& x = t == 0 ? x0 : x1;
          uint8_t* q = sQ + (size_t)(qb * 2 + t) * TILE_BYTES;
          if constexpr (MODE == MODE_GLOBAL)
            tc::tma_load_3d(q, &tmap, &bars->q_full[qb], x.head * DH, x.q0, x.b);
          else if constexpr (MODE == MODE_NA)
            tc::tma_load_4d(q, &tmap, &bars->q_full[qb], x.head * DH, x.qj0, x.qi0, x.b);
          else
            load_window_tile(q, 0, x, &bars->q_full[qb]);
        }
        for (int j = 0; j < nb; ++j) {
          kv_stage(x0, j);
          if constexpr (!SHARED_KV) kv_stage(x1, j);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (tc::elect_one()) {
      constexpr uint32_t IDESC_S = tc::idesc_bf16(ROWS, 128);
      constexpr uint32_t IDESC_O = tc::idesc_bf16(ROWS, DH, 0, 1);
      const uint32_t q_base = tc::smem_u32(sQ), kv_base = tc::smem_u32(sKV), p_base = tc::smem_u32(sP);
      uint32_t rs = 0, rph = 0;         // ring cursor: next K/V stage in consumption order
      uint32_t n_p[2] = {0, 0};         // P tiles consumed per group (parity of p_ready)
      auto take = [&](uint32_t& s, uint32_t& ph) {       // hand out the cursor's stage, advance the cursor
        s = rs;
        ph = rph;
        if (++rs == PA_STAGES) {
          rs = 0;
          rph ^= 1u;
        }
      };
      auto issue_s = [&](int t, int qb, uint32_t stage) {
        const uint64_t qd = tc::smem_desc_k_sw128(q_base + (uint32_t)((qb * 2 + t) * TILE_BYTES));
        const uint64_t kd = tc::smem_desc_k_sw128(kv_base + (uint32_t)((stage * 2) * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc::umma_bf16(tmem + (uint32_t)(t * 128), qd + 2ull * k, kd + 2ull * k, IDESC_S, (uint32_t)(k != 0));
        tc::umma_commit(&bars->s_ready[t]);
        PA_TRACE(1, 10 + t);
      };
      for (int u = 0; u < n_local; ++u) {
        const int qb = u & 1;
        uint32_t cs[2], cph[2], ns[2] = {0, 0}, nph[2] = {0, 0};
        take(cs[0], cph[0]);
        if constexpr (SHARED_KV) {
          cs[1] = cs[0];
          cph[1] = cph[0];
        } else {
          take(cs[1], cph[1]);
        }
        tc::mbar_wait(&bars->q_full[qb], (uint32_t)((u >> 1) & 1));
        PA_TRACE(1, 2);
        // S tiles of the pair's first key block.  S_t is free: p_ready of the previous pair's last block was waited below.
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t == 0 || !SHARED_KV) tc::mbar_wait(&bars->kv_full[cs[t]], cph[t]);
          tc::tc_fence_after();
          issue_s(t, qb, cs[t]);
        }
        for (int j = 0; j < nb; ++j) {
          const bool more = j + 1 < nb;
          if (more) {
            take(ns[0], nph[0]);
            if constexpr (SHARED_KV) {
              ns[1] = ns[0];
              nph[1] = nph[0];
            } else {
              take(ns[1], nph[1]);
            }
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            tc::mbar_wait(&bars->p_ready[t], n_p[t] & 1u);          // P_t(j) is in shared memory, S_t(j) fully read
            ++n_p[t];
            PA_TRACE(1, 20 + t);
            // (O_t of the previous pair has left TMEM: the group reads it before it starts the tile whose P_t was just awaited)
            tc::tc_fence_after();
            const uint64_t vd = tc::smem_desc_mn_sw128(kv_base + (uint32_t)((cs[t] * 2 + 1) * TILE_BYTES), 1024, 1024);
            if constexpr (PT) {        // A = P_t in TMEM: 8 columns (16 bf16) per k-step, over the first 64 columns of S_t
#pragma unroll
              for (int k = 0; k < ROWS / 16; ++k)
                tc::umma_bf16_ts(tmem + 256u + (uint32_t)(t * 64), tmem + (uint32_t)(t * 128 + k * 8), vd + (uint64_t)(k * ((16 * 128) >> 4)), IDESC_O,
                                 (uint32_t)((j | k) != 0));
            } else {
              const uint64_t pd = tc::smem_desc_k_sw128(p_base + (uint32_t)(t * 2 * TILE_BYTES));
#pragma unroll
              for (int k = 0; k < ROWS / 16; ++k)
                tc::umma_bf16(tmem + 256u + (uint32_t)(t * 64), pd + (uint64_t)((k >> 2) * (TILE_BYTES >> 4)) + 2ull * (k & 3),
                              vd + (uint64_t)(k * ((16 * 128) >> 4)), IDESC_O, (uint32_t)((j | k) != 0));
            }
            tc::umma_commit(&bars->pv_done[t]);
            if (!SHARED_KV || t == 1) tc::umma_commit(&bars->kv_empty[cs[t]]);      // every MMA that reads this stage has been issued
            if (more) {                                                             // next block's S tile for the group that just finished
              if (t == 0 || !SHARED_KV) tc::mbar_wait(&bars->kv_full[ns[t]], nph[t]);
              tc::tc_fence_after();
              issue_s(t, qb, ns[t]);
            }
          }
          cs[0] = ns[0];
          cs[1] = ns[1];
          cph[0] = nph[0];
          cph[1] = nph[1];
        }
        tc::umma_commit(&bars->q_empty[qb]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups: thread = query row
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem + (uint32_t)(t * 128) + lane_base, tmem_o = tmem + 256u + (uint32_t)(t * 64) + lane_base;
    uint8_t* const pt = sP + (size_t)t * 2 * TILE_BYTES;
    const int hd = row >> 6, quad = (row & 63) >> 4;        // WINDOW: head of this row, its 4x4 quadrant in the (rolled) window
    uint32_t n = 0;                    // tiles processed by this group (parity of s_ready / pv_done)
    for (int u = 0; u < n_local; ++u) {
      const int pair = (int)blockIdx.x + u * (int)gridDim.x;
      const PipeTile x = pipe_decode<MODE>(p, pair, t);
      const float mb = __ldg(p.bound + x.head + (MODE == MODE_WINDOW ? hd : 0)) * LOG2E;
      // NA: this row's query and the origin of its clamped 7x7 window, relative to the halo
      int na_a = 0, na_b = 0, na_qi = 0, na_qj = 0;
      if constexpr (MODE == MODE_NA) {
        na_qi = x.qi0 + (row >> 4);
        na_qj = x.qj0 + (row & 15);
        na_a = min(max(na_qi - 3, 0), p.h - 7) - x.r0;         // 0..7: first halo row of the neighbourhood
        na_b = min(max(na_qj - 3, 0), p.w - 7) - x.c0;         // 0..15: first halo column
      }
      uint32_t wmask[2] = {0xffffffffu, 0xffffffffu};            // WINDOW: seam mask of the two 32-key chunks of the own head
      if constexpr (MODE == MODE_WINDOW) {
        const bool seam_r = p.shift > 0 && x.qi0 == 0, seam_c = p.shift > 0 && x.qj0 == 0;
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
          const bool ok = (!seam_r || ((kq >> 1) == (quad >> 1))) && (!seam_c || ((kq & 1) == (quad & 1)));
          if (!ok) wmask[kq >> 1] &= (kq & 1) ? 0x0000ffffu : 0xffff0000u;
        }
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      for (int j = 0; j < nb; ++j, ++n) {
        // NA: 128-bit key mask of this (row, block): bit (dr * 22 + dc) for halo row 5 j + dr inside [a, a + 7), dc inside [b, b + 7)
        uint32_t km[4] = {0u, 0u, 0u, 0u};
        if constexpr (MODE == MODE_NA) {
          const unsigned long long rowbits = 0x7Full << na_b;
          unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll
          for (int dr = 0; dr < NA_BLK_ROWS; ++dr) {
            const int hr = j * NA_BLK_ROWS + dr;
            if ((unsigned)(hr - na_a) < 7u) {
              constexpr int W = NA_KW;
              if (dr * W < 64) lo |= rowbits << (dr * W);
              if (dr * W < 64 && dr * W + W > 64) hi |= rowbits >> (64 - dr * W);
              if (dr * W >= 64) hi |= rowbits << (dr * W - 64);
            }
          }
          km[0] = (uint32_t)lo;
          km[1] = (uint32_t)(lo >> 32);
          km[2] = (uint32_t)hi;
          km[3] = (uint32_t)(hi >> 32);
        }
        tc::mbar_wait(&bars->s_ready[t], n & 1u);
        tc::tc_fence_after();
        if (row == 0) PA_TRACE(2 + t, 30);
        // which of each chunk's 32 keys count for this row, and whether the warp needs the chunk at all (warp-uniform)
        uint32_t mw[4];
        bool need[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mw[c] = 0xffffffffu;
          need[c] = true;
          if constexpr (MODE == MODE_NA) {
            mw[c] = km[c];
            need[c] = __any_sync(0xffffffffu, mw[c] != 0u);
          } else if constexpr (MODE == MODE_WINDOW) {
            need[c] = (c >> 1) == hd;
            mw[c] = wmask[c & 1];
          }
        }
        // (the P tile is free: s_ready of this block was committed after the previous block's P V MMAs, so they have completed)
        // S chunks are double buffered in registers: tcgen05.wait::ld waits for EVERY outstanding load, so the load of chunk c + 1
        // is issued right after the wait for chunk c and streams in behind chunk c's exponentials
        uint32_t r[2][32];
        if (need[0]) tc::tmem_ld32_nowait(tmem_s, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (need[c]) tc::tmem_ld_wait(r[c & 1]);
          if (c < 3 && need[c < 3 ? c + 1 : 3]) tc::tmem_ld32_nowait(tmem_s + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
          if (MODE == MODE_WINDOW && !PT && !need[c]) continue;        // off-diagonal half: zero since the prologue
          uint32_t pk[16];
          if (need[c]) {
            const uint32_t* v = r[c & 1];
            const uint32_t mwc = mw[c];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float e[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                e[q] = tc::ex2(fmaf(__uint_as_float(v[2 * i + q]), LOG2E, -mb));
                if constexpr (MODE != MODE_GLOBAL) e[q] = (mwc & (1u << (2 * i + q))) ? e[q] : 0.f;
              }
              l0 += e[0];
              l1 += e[1];
              l2 += e[2];
              l3 += e[3];
              pk[i] = tc::pack_bf16x2(e[0], e[1]);
              pk[i + 1] = tc::pack_bf16x2(e[2], e[3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          if constexpr (PT) {          // columns [16 c, 16 c + 16) of the own S tile: already read (chunk c covers columns up to 32 c + 31)
            tc::tmem_st16(tmem_s + (uint32_t)(c * 16), pk);
          } else {
            uint8_t* ph_ = pt + (c >> 1) * TILE_BYTES;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              *reinterpret_cast<uint4*>(ph_ + p_offset(row, (c & 1) * 4 + jj)) = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
          }
        }
        if constexpr (PT)
          tc::tmem_st_wait();
        else
          tc::fence_proxy_async();
        tc::tc_fence_before();
        tc::mbar_arrive(&bars->p_ready[t]);
        if (row == 0) PA_TRACE(2 + t, 31);
      }
      // ---- O_t / l -> out
      tc::mbar_wait(&bars->pv_done[t], (n - 1) & 1u);
      tc::tc_fence_after();
      uint32_t o0[32], o1[32];
      tc::tmem_ld32_nowait(tmem_o, o0);
      tc::tmem_ld32_nowait(tmem_o + 32, o1);
      tc::tmem_ld_wait(o0);
      tc::tmem_ld_wait(o1);
      tc::tc_fence_before();
      if (row == 0) PA_TRACE(2 + t, 32);
      const float inv = 1.f / ((l0 + l1) + (l2 + l3));
      int64_t token;
      int head = x.head;
      if constexpr (MODE == MODE_GLOBAL) {
        token = (int64_t)x.q0 + row;
      } else if constexpr (MODE == MODE_NA) {
        token = (int64_t)na_qi * p.w + na_qj;
      } else {
        int oi, oj;
        if (p.shift == 0) {
          oi = x.qi0 * 8 + ((row & 63) >> 3);
          oj = x.qj0 * 8 + (row & 7);
        } else {
          const int lr = (row & 15) >> 2, lc = row & 3;
          oi = (x.qi0 * 8 + (quad >> 1) * 4 + lr - p.shift + p.h) % p.h;
          oj = (x.qj0 * 8 + (quad & 1) * 4 + lc - p.shift + p.w) % p.w;
        }
        token = (int64_t)oi * p.w + oj;
        head = x.head + hd;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.out + (((int64_t)x.b * p.h * p.w + token) * nh + head) * DH);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        dst[jj] = make_uint4(tc::pack_bf16x2(__uint_as_float(o0[jj * 8 + 0]) * inv, __uint_as_float(o0[jj * 8 + 1]) * inv),
                             tc::pack_bf16x2(__uint_as_float(o0[jj * 8 + 2]) * inv, __uint_as_float(o0[jj * 8 + 3]) * inv),
                             tc::pack_bf16x2(__uint_as_float(o0[jj * 8 + 4]) * inv, __uint_as_float(o0[jj * 8 + 5]) * inv),
                             tc::pack_bf16x2(__uint_as_float(o0[jj * 8 + 6]) * inv, __uint_as_float(o0[jj * 8 + 7]) * inv));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        dst[4 + jj] = make_uint4(tc::pack_bf16x2(__uint_as_float(o1[jj * 8 + 0]) * inv, __uint_as_float(o1[jj * 8 + 1]) * inv),
                                 tc::pack_bf16x2(__uint_as_float(o1[jj * 8 + 2]) * inv, __uint_as_float(o1[jj * 8 + 3]) * inv),
                                 tc::pack_bf16x2(__uint_as_float(o1[jj * 8 + 4]) * inv, __uint_as_float(o1[jj * 8 + 5]) * inv),
                                 tc::pack_bf16x2(__uint_as_float(o1[jj * 8 + 6]) * inv, __uint_as_float(o1[jj * 8 + 7]) * inv));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

// host: launch the pipelined kernel
template <int MODE, bool PT>
static int launch_attn_pipe_impl(const CUtensorMap& tq, const CUtensorMap& tkv, PipeAttnParams p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    KDB_CUDA(cudaFuncSetAttribute(attn_pipe_kernel<MODE, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PA_SMEM));
    attr_set = true;
  }
  static const bool no_pdl = [] {
    const char* e_ = getenv("KDB200_NO_PDL");
    return e_ != nullptr && e_[0] == '1';
  }();
  static const bool trace_on = [] {
    const char* e_ = getenv("KDB200_ATTN_TRACE");
    return e_ != nullptr && e_[0] == '1';
  }();
  static long long* trace_buf = nullptr;
  if (trace_on) {
    if (trace_buf == nullptr) KDB_CUDA(cudaMalloc(&trace_buf, 4 * PA_TRACE_N * 2 * sizeof(long long)));
    KDB_CUDA(cudaMemsetAsync(trace_buf, 0, 4 * PA_TRACE_N * 2 * sizeof(long long), st));
    p.trace = trace_buf;
  }
  int sms = 0, dev = 0;
  KDB_CUDA(cudaGetDevice(&dev));
  KDB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)(p.n_pairs < sms ? p.n_pairs : sms));
  lc.blockDim = dim3(320);
  lc.dynamicSmemBytes = PA_SMEM;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = no_pdl || trace_on ? 0 : 1;
  KDB_CUDA(cudaLaunchKernelEx(&lc, attn_pipe_kernel<MODE, PT>, tq, tkv, p));
  if (trace_on) {
    static long long hbuf[4 * PA_TRACE_N * 2];
    KDB_CUDA(cudaMemcpyAsync(hbuf, trace_buf, sizeof(hbuf), cudaMemcpyDeviceToHost, st));
    KDB_CUDA(cudaStreamSynchronize(st));
    long long t0 = 0;
    for (int r = 0; r < 4; ++r)
      if (hbuf[(r * PA_TRACE_N) * 2 + 1] != 0 && (t0 == 0 || hbuf[(r * PA_TRACE_N) * 2 + 1] < t0)) t0 = hbuf[(r * PA_TRACE_N) * 2 + 1];
    fprintf(stderr, "ATTN trace MODE=%d PT=%d B=%d h=%d w=%d nh=%d nb=%d pairs=%d grid=%u (cycles since the first stamp of CTA 0; codes: 1 Q issue, 100+j K/V issue, 2 q_full, 10+t S issued, 20+t P ready seen, 30 S ready seen, 31 P written, 32 O read)\n",
            MODE, (int)PT, p.B, p.h, p.w, p.nh, p.nb, p.n_pairs, lc.gridDim.x);
    const char* names[4] = {"producer", "mma", "softmax0", "softmax1"};
    for (int r = 0; r < 4; ++r) {
      fprintf(stderr, " %-9s", names[r]);
      for (int i = 0; i < PA_TRACE_N && hbuf[(r * PA_TRACE_N + i) * 2 + 1] != 0; ++i)
        fprintf(stderr, " %lld@%lld", hbuf[(r * PA_TRACE_N + i) * 2], hbuf[(r * PA_TRACE_N + i) * 2 + 1] - t0);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

template <int MODE>
static int launch_attn_pipe(const CUtensorMap& tq, const CUtensorMap& tkv, const PipeAttnParams& p, cudaStream_t st) {
  static const bool p_smem = [] {
    const char* e_ = getenv("KDB200_ATTN_P_SMEM");
    return e_ != nullptr && e_[0] == '1';
  }();
  return p_smem ? launch_attn_pipe_impl<MODE, false>(tq, tkv, p, st) : launch_attn_pipe_impl<MODE, true>(tq, tkv, p, st);
}
