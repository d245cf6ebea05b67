// tc_attention_persist.cuh -- EXPERIMENTAL persistent shifted-window attention (included inside tc_attention.cu's anonymous
// namespace; selected only with KDB200_ATTN_PERSIST=1, default off; first GPU run pending -- see tests/test_gpu_zz_experimental.py).
//
// Motivation (profiles/r1_ncu_full_summary.json): attn_tc_kernel<WINDOW> runs one (window, head pair) per CTA as a
// serial chain TMA -> S MMA -> softmax -> O MMA -> store, four CTAs per SM; 31 % of a softmax warp's life is the wait for
// Q/K/V and the kernel moves 134 MB in 41-48 us (3 TB/s).  Here a CTA loops over units with TWO shared-memory buffers and TWO
// TMEM accumulators: the loads of unit i+2 are issued as soon as the P V MMA of unit i has drained buffer i & 1, and S of
// unit i+1 is issued while the softmax threads still work on unit i.  Arithmetic is the same code as the one-shot kernel, so
// results must be bit-identical to it.
//
//   unit u = blockIdx.x + i * gridDim.x  ->  (image, head pair, window row, window column)
//   buffer b = i & 1: sQ | sK | sV (3 x 16 KiB), P overwrites Q,K once S has completed; TMEM columns [128 b, 128 b + 128): S, O on its first 64
//   barriers per buffer, use index i >> 1:  full (TMA bytes)  s_ready (S MMA done)  p_ready (128 softmax threads wrote P)
//                                           o_ready (P V MMA done: O valid AND the buffer may be refilled)  o_consumed (O is in registers)

struct PersistAttnBars {
  uint64_t full[2], s_ready[2], p_ready[2], o_ready[2], o_consumed[2];
  uint32_t tmem;
};

struct WindowUnit {
  int b, head0, wi, wj;
};

__device__ __forceinline__ WindowUnit decode_unit(int u, const AttnParams& p) {
  const int nww = p.w / 8, nwh = p.h / 8, hp = p.nh / 2;
  WindowUnit r;
  r.wj = u % nww;
  u /= nww;
  r.wi = u % nwh;
  u /= nwh;
  r.head0 = (u % hp) * 2;
  r.b = u / hp;
  return r;
}

__global__ void __launch_bounds__(160, 2) attn_window_persist_kernel(const __grid_constant__ CUtensorMap tmap, const AttnParams p, const int n_units) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  PersistAttnBars* bars = reinterpret_cast<PersistAttnBars*>(base + 6 * TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = p.nh;
  const int n_local = (int)blockIdx.x < n_units ? (n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap);
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&bars->full[b], 1);
      tc::mbar_init(&bars->s_ready[b], 1);
      tc::mbar_init(&bars->p_ready[b], 128);
      tc::mbar_init(&bars->o_ready[b], 1);
      tc::mbar_init(&bars->o_consumed[b], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 4) tc::tmem_alloc(&bars->tmem, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  tc::pdl_wait();
  tc::pdl_launch_dependents();
  const uint32_t tmem = bars->tmem;

  if (warp == 4) {
    if (tc::elect_one()) {
      auto load_tile = [&](uint8_t* dst, int t, const WindowUnit& u, uint64_t* bar) {
        if (p.shift == 0) {   // unshifted window: one 8x8 box per head (tensor map box = 64 x 8 x 8)
#pragma unroll
          for (int hd = 0; hd < 2; ++hd) tc::tma_load_4d(dst + hd * 64 * 128, &tmap, bar, (t * nh + u.head0 + hd) * DH, u.wj * 8, u.wi * 8, u.b);
          return;
        }
#pragma unroll
        for (int hd = 0; hd < 2; ++hd)
#pragma unroll
          for (int quad = 0; quad < 4; ++quad) {
            const int r0 = (u.wi * 8 + (quad >> 1) * 4 - p.shift + p.h) % p.h;   // rolled -> original coordinates (:274)
            const int c0 = (u.wj * 8 + (quad & 1) * 4 - p.shift + p.w) % p.w;
            tc::tma_load_4d(dst + (hd * 64 + quad * 16) * 128, &tmap, bar, (t * nh + u.head0 + hd) * DH, c0, r0, u.b);
          }
      };
      auto load_unit = [&](int i) {
        const int b = i & 1;
        const WindowUnit u = decode_unit((int)blockIdx.x + i * (int)gridDim.x, p);
        uint8_t* buf = base + b * 3 * TILE_BYTES;
        tc::mbar_arrive_expect_tx(&bars->full[b], 3 * TILE_BYTES);
        load_tile(buf, 0, u, &bars->full[b]);
        load_tile(buf + TILE_BYTES, 1, u, &bars->full[b]);
        load_tile(buf + 2 * TILE_BYTES, 2, u, &bars->full[b]);
      };
      constexpr uint32_t IDESC_S = tc::idesc_bf16(ROWS, 128);
      constexpr uint32_t IDESC_O = tc::idesc_bf16(ROWS, DH, 0, 1);
      auto issue_s = [&](int i) {        // S = [Q_h0; Q_h1] [K_h0; K_h1]^T of unit i into TMEM buffer i & 1
        const int b = i & 1;
        const uint32_t sq = tc::smem_u32(base + b * 3 * TILE_BYTES);
        const uint64_t qdesc = tc::smem_desc_k_sw128(sq), kdesc = tc::smem_desc_k_sw128(sq + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc::umma_bf16(tmem + b * 128, qdesc + 2ull * k, kdesc + 2ull * k, IDESC_S, (uint32_t)(k != 0));
        tc::umma_commit(&bars->s_ready[b]);
      };
      if (n_local > 0) load_unit(0);
      if (n_local > 1) load_unit(1);
      if (n_local > 0) {
        tc::mbar_wait(&bars->full[0], 0);
        tc::tc_fence_after();
        issue_s(0);
      }
      for (int i = 0; i < n_local; ++i) {
        const int b = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        if (i + 1 < n_local) {                      // S of the next unit while the softmax threads work on this one
          const int b1 = b ^ 1;
          tc::mbar_wait(&bars->full[b1], (uint32_t)((i + 1) >> 1) & 1u);
          if (i >= 1) tc::mbar_wait(&bars->o_consumed[b1], (uint32_t)((i - 1) >> 1) & 1u);   // O of unit i-1 has left that TMEM buffer
          tc::tc_fence_after();
          issue_s(i + 1);
        }
        tc::mbar_wait(&bars->p_ready[b], par);      // P written, S consumed
        tc::tc_fence_after();
        {
          const uint32_t sp = tc::smem_u32(base + b * 3 * TILE_BYTES);          // P overwrites Q | K
          const uint64_t pdesc = tc::smem_desc_k_sw128(sp);
          const uint64_t vdesc = tc::smem_desc_mn_sw128(sp + 2 * TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < ROWS / 16; ++k) {
            const uint64_t ad = pdesc + (uint64_t)((k >> 2) * (TILE_BYTES >> 4)) + 2ull * (k & 3);
            const uint64_t bd = vdesc + (uint64_t)(k * ((16 * 128) >> 4));
            tc::umma_bf16(tmem + b * 128, ad, bd, IDESC_O, (uint32_t)(k != 0));
          }
          tc::umma_commit(&bars->o_ready[b]);
        }
        if (i + 2 < n_local) {                      // the P V MMA has drained this buffer: refill it
          tc::mbar_wait(&bars->o_ready[b], par);
          load_unit(i + 2);
        }
      }
    }
  } else {
    // ---------------------------------------------------- softmax / epilogue: thread = row (same arithmetic as attn_tc_kernel<WINDOW>)
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int hd = row >> 6, quad = (row & 63) >> 4;
    for (int i = 0; i < n_local; ++i) {
      const int b = i & 1;
      const uint32_t par = (uint32_t)(i >> 1) & 1u;
      const WindowUnit u = decode_unit((int)blockIdx.x + i * (int)gridDim.x, p);
      const bool seam_r = p.shift > 0 && u.wi == 0, seam_c = p.shift > 0 && u.wj == 0;
      uint8_t* sP = base + b * 3 * TILE_BYTES;
      const uint32_t tmem_s = tmem + b * 128;
      float m = -INFINITY, l = 0.f;
      tc::mbar_wait(&bars->s_ready[b], par);
      tc::tc_fence_after();
      {
        float v[64];
        {
          float t0[32], t1[32];
          tc::tmem_ld32(tmem_s + lane_base + hd * 64, t0);
          tc::tmem_ld32(tmem_s + lane_base + hd * 64 + 32, t1);
#pragma unroll
          for (int k = 0; k < 32; ++k) { v[k] = t0[k]; v[32 + k] = t1[k]; }
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const int kq = k >> 4;
          const bool ok = (!seam_r || ((kq >> 1) == (quad >> 1))) && (!seam_c || ((kq & 1) == (quad & 1)));
          v[k] = ok ? v[k] : -INFINITY;
          m = fmaxf(m, v[k]);
        }
        const float mb = m * LOG2E;
        uint8_t* own = sP + hd * TILE_BYTES;
        uint8_t* other = sP + (1 - hd) * TILE_BYTES;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float p0 = exp2f(fmaf(v[jj * 8 + 2 * t], LOG2E, -mb)), p1 = exp2f(fmaf(v[jj * 8 + 2 * t + 1], LOG2E, -mb));
            pk[t] = tc::pack_bf16x2(p0, p1);
            float q0, q1;
            tc::unpack_bf16x2(pk[t], q0, q1);
            l += q0 + q1;
          }
          *reinterpret_cast<uint4*>(own + p_offset(row, jj)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(other + p_offset(row, jj)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(&bars->p_ready[b]);
      // ------------------------------------------------------ O / l -> out
      tc::mbar_wait(&bars->o_ready[b], par);
      tc::tc_fence_after();
      float o0[32], o1[32];
      tc::tmem_ld32(tmem_s + lane_base, o0);
      tc::tmem_ld32(tmem_s + lane_base + 32, o1);
      tc::tc_fence_before();
      tc::mbar_arrive(&bars->o_consumed[b]);        // this TMEM buffer may take the S of unit i + 2
      const float inv = 1.f / l;
      int oi, oj;
      if (p.shift == 0) {
        oi = u.wi * 8 + ((row & 63) >> 3);
        oj = u.wj * 8 + (row & 7);
      } else {
        const int lr = (row & 15) >> 2, lc = row & 3;
        oi = (u.wi * 8 + (quad >> 1) * 4 + lr - p.shift + p.h) % p.h;
        oj = (u.wj * 8 + (quad & 1) * 4 + lc - p.shift + p.w) % p.w;
      }
      const int64_t token = (int64_t)oi * p.w + oj;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (((int64_t)u.b * p.h * p.w + token) * nh + (u.head0 + hd)) * DH);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        dst[jj] = make_uint4(tc::pack_bf16x2(o0[jj * 8 + 0] * inv, o0[jj * 8 + 1] * inv), tc::pack_bf16x2(o0[jj * 8 + 2] * inv, o0[jj * 8 + 3] * inv),
                             tc::pack_bf16x2(o0[jj * 8 + 4] * inv, o0[jj * 8 + 5] * inv), tc::pack_bf16x2(o0[jj * 8 + 6] * inv, o0[jj * 8 + 7] * inv));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        dst[4 + jj] = make_uint4(tc::pack_bf16x2(o1[jj * 8 + 0] * inv, o1[jj * 8 + 1] * inv), tc::pack_bf16x2(o1[jj * 8 + 2] * inv, o1[jj * 8 + 3] * inv),
                                 tc::pack_bf16x2(o1[jj * 8 + 4] * inv, o1[jj * 8 + 5] * inv), tc::pack_bf16x2(o1[jj * 8 + 6] * inv, o1[jj * 8 + 7] * inv));
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 256);
  }
}

constexpr size_t ATTN_SMEM_WINDOW_PERSIST = 6 * TILE_BYTES + 1024 + 256;
