// tc_gemm_persist.cuh -- persistent, warp-specialised tcgen05 GEMM (included inside tc_kernels.cu's anonymous namespace).
//
// One CTA per SM (320 threads) walks a list of 128 x 128 output tiles:
//   warp 0      TMA producer: a ring of SWIZZLE_128B stages that runs ahead across tiles.
//   warp 1      MMA issuer: tcgen05.mma (M=128, N=128, K=16) into one of TWO TMEM accumulators (2 x 128 columns); a commit frees
//               the ring stage, the commit after the last k-block signals tmem_full[acc].
//   warps 2-9   two ping-pong epilogue groups of four warps; group g owns accumulator g, staging tile(s) g and every tile with
//               (it & 1) == g.  Thread = accumulator row: tcgen05.ld (two passes of 64 columns), release the accumulator as soon
//               as it is in registers, fused epilogue in fp32, bf16 pack into a swizzled staging tile, one thread issues the
//               TMA store.  Residual / skip tiles are loaded by TMA INTO the staging tile and added in place.
// Steady state per tile = max(TMA, MMA, epilogue) instead of their sum; measured, the epilogue groups set the pace for K <= 256
// (see DESIGN.md section 4), so everything that is not arithmetic is kept off their critical path: waits for the previous TMA
// store are deferred behind pass 0, table rows are prefetched, barrier waits are parked in hardware (suspend-time hint).
//
// Weight-resident mode (small K): a CTA owns 1-3 n-blocks whose [128 x K] weight blocks stay in shared memory for the whole
// kernel; it walks m-tiles and uses ONE load of each A tile for all of its n-blocks.  The grid is rounded to a multiple of the
// number of n-block groups.  Otherwise (streaming mode) both operands go through the ring, tile = blockIdx.x + it * gridDim.x.
//
// The kernel is launched programmatically (cudaLaunchAttributeProgrammaticStreamSerialization): barrier init, TMEM allocation and
// the resident weight load happen before griddepcontrol.wait, i.e. while the previous kernel in the stream is still draining.
#pragma once

constexpr int P_BN = 128;
// Epilogue groups (4 warps each, one TMEM accumulator + staging tile per group).  The GEGLU epilogue is the heaviest (64 MUFU.TANH
// + ~220 packed FMA-pipe instructions per thread and tile against 512 cycles of tensor work) and ncu shows its warps stalled on
// fixed-latency dependencies with 37 % of the issue slots used: a THIRD group gives the schedulers another warp each.  Its 146
// registers x 448 threads just fit the register file; the other epilogues (162-166 registers) stay at two groups.
template <int EPI> constexpr int epi_groups() { return EPI == TCE_GEGLU ? 3 : 2; }
template <int EPI> constexpr int persist_threads() { return 96 + 128 * epi_groups<EPI>(); }      // producer + two MMA issuers + the epilogue groups
constexpr int P_MAX_GROUPS = 3;
constexpr int P_B_TILE_BYTES = P_BN * BK * 2;                     // 16 KiB: one k-block of the weight block
constexpr int P_OUT_BYTES = 2 * SUB_TILE_BYTES;                   // 128 x 128 bf16 staging tile
constexpr int P_MAX_STAGES = 8;
constexpr size_t P_SMEM_LIMIT = 227 * 1024;

struct PersistBars {
  uint64_t full[P_MAX_STAGES], empty[P_MAX_STAGES];
  uint64_t tmem_full[P_MAX_GROUPS], tmem_empty[P_MAX_GROUPS];
  uint64_t resid_full[2][2], b_full;   // resid_full[group][buffer]: a waiter must observe every phase of its barrier
  uint64_t turn[2];                    // MMA issuers: turn[i] = issuer i may start the waits of its next tile (see the MMA role)
  uint32_t tmem;
};

// debug timestamps: CTA 0, first 32 tiles, 16 slots per tile
#define KDB_TRACE(slot_)                                                                          \
  do {                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && it < 32) p.trace[it * 16 + (slot_)] = clock64(); \
  } while (0)

struct PersistCfg {
  int stages, b_res, sc_bufs;
  int nb;   // weight-resident mode: n-blocks kept resident per CTA; every A tile is loaded once and used for all of them
  int issuers;    // 1 or 2 MMA-issuing threads (KDB200_GEMM_ISSUERS); see the MMA role below
};

template <int EPI>
__global__ void __launch_bounds__(persist_threads<EPI>(), 1) gemm_tc_persist(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
                                                                const __grid_constant__ CUtensorMap tmc, const __grid_constant__ CUtensorMap tmr,
                                                                const TcParams p, const PersistCfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t IDESC = tc::idesc_bf16(BM, P_BN);
  // RESID / SPLIT: the [128 x 128] residual / skip tile is loaded by TMA straight INTO the group's output staging tile and the
  // epilogue adds in place, so it is double-buffered with the staging tiles and needs no shared memory of its own.
  constexpr bool RES = EPI == TCE_RESID || EPI == TCE_SPLIT;
  constexpr int NG = epi_groups<EPI>();
  constexpr uint32_t TMEM_COLS = NG > 2 ? 512 : 256;
  // GEGLU writes 64 output columns per 128 accumulator columns: its staging tile is one [128 x 64] sub-tile
  constexpr int OUT_BYTES = EPI == TCE_GEGLU ? SUB_TILE_BYTES : P_OUT_BYTES;
  const int nkb = p.K / BK;
  const int stage_bytes = cfg.b_res ? A_STAGE_BYTES : A_STAGE_BYTES + P_B_TILE_BYTES;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                                                            // resident weight block (b_res)
  const int nb = cfg.b_res ? cfg.nb : 1;
  uint8_t* sStage = sB + (cfg.b_res ? (size_t)nb * nkb * P_B_TILE_BYTES : 0);
  uint8_t* sC = sStage + (size_t)cfg.stages * stage_bytes;                       // 1 or 2 staging tiles
  PersistBars* bars = reinterpret_cast<PersistBars*>(sC + (size_t)cfg.sc_bufs * OUT_BYTES);

  // Warp roles.  The SMSP arbiter picks the eligible warp with the HIGHEST warp id first (B300_MICROARCH.md, "multi-warp
  // arbiter"): the MMA-issuing threads must never queue behind epilogue warps that are always eligible, so the producer and the
  // two MMA issuers are the three highest warps of the CTA and the epilogue groups are warps 0 .. 4 NG - 1.  `warp` below is the
  // ROLE index (0 producer, 1 / 2 MMA issuer 0 / 1, 3.. epilogue); TMEM lane quadrants use the physical warp id.
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = pwarp >= 4 * NG ? pwarp - 4 * NG : pwarp + 3;
  const int n_tiles_n = p.N / P_BN;
  const int m_tiles = (int)((p.M + BM - 1) / BM);
  const int n_tiles = n_tiles_n * m_tiles;
  // Work list of this CTA, as a sequence of 128 x 128 output tiles `it` = 0 .. n_local-1.
  //  streaming mode : tile = blockIdx.x + it * gridDim.x over the (m, n) grid, both operands loaded per tile
  //  resident mode  : the CTA owns n-blocks [ng*nb, ng*nb+nb) (weights resident) and walks m-tiles mw, mw+G, ...; the nb tiles of
  //                   one m-tile are consecutive and share ONE load of the A tile (A leaves L2 once per nb*128 output columns)
  int n_local, ng = 0, mw = 0, G = 1;
  if (cfg.b_res) {
    const int n_groups = n_tiles_n / nb;
    ng = (int)blockIdx.x % n_groups;
    mw = (int)blockIdx.x / n_groups;
    G = (int)gridDim.x / n_groups;
    n_local = mw < m_tiles ? ((m_tiles - 1 - mw) / G + 1) * nb : 0;
  } else {
    n_local = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  }
  auto coords = [&](uint32_t it, int& m0, int& n0, int& j) {
    if (cfg.b_res) {
      const int i = (int)it / nb;
      j = (int)it - i * nb;
      m0 = (mw + i * G) * BM;
      n0 = (ng * nb + j) * P_BN;
    } else {
      const int tile = (int)blockIdx.x + (int)it * (int)gridDim.x;
      j = 0;
      m0 = (tile / n_tiles_n) * BM;
      n0 = (tile % n_tiles_n) * P_BN;
    }
  };

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma);
    tc::tma_prefetch_desc(&tmb);
    tc::tma_prefetch_desc(&tmc);
    if (RES) tc::tma_prefetch_desc(&tmr);
    for (int s = 0; s < cfg.stages; ++s) {
      tc::mbar_init(&bars->full[s], 1);
      tc::mbar_init(&bars->empty[s], (cfg.b_res && nb >= 2 && cfg.issuers == 2) ? 2u : 1u);   // one commit per issuer that reads the stage
    }
    for (int a = 0; a < NG; ++a) {
      tc::mbar_init(&bars->tmem_full[a], 1);
      tc::mbar_init(&bars->tmem_empty[a], 128);          // one ping-pong epilogue group (4 warps)
    }
    for (int a = 0; a < 4; ++a) tc::mbar_init(&bars->resid_full[a >> 1][a & 1], 1);
    tc::mbar_init(&bars->b_full, 1);
    tc::mbar_init(&bars->turn[0], 1);
    tc::mbar_init(&bars->turn[1], 1);
    tc::fence_barrier_init();
    // The resident weight block does not depend on the previous kernel: it is requested before the programmatic-launch wait
    // below, so (with the prologue) it overlaps the tail of the predecessor.
    if (cfg.b_res && n_local > 0) {
      tc::mbar_arrive_expect_tx(&bars->b_full, (uint32_t)(nb * nkb) * P_B_TILE_BYTES);
      for (int j = 0; j < nb; ++j)
        for (int kb = 0; kb < nkb; ++kb)
          tc::tma_load_2d(sB + (size_t)(j * nkb + kb) * P_B_TILE_BYTES, &tmb, &bars->b_full, kb * BK, (ng * nb + j) * P_BN);
    }
  }
  if (warp == 1) tc::tmem_alloc(&bars->tmem, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = bars->tmem;
  tc::pdl_wait();                    // everything below reads or writes tensors the previous kernel may still be producing
  tc::pdl_launch_dependents();

  if (warp == 0) {
    if (tc::elect_one() && !(p.dbg & 32)) {
      uint32_t s = 0, ph = 0;      // ring slot / phase (carried incrementally: no division in the loop)
      int jj = 0;
      for (uint32_t it = 0; it < (uint32_t)n_local; ++it) {
        const bool load = jj == 0;          // otherwise the A tile of this m-tile is already in the ring
        if (++jj == nb) jj = 0;
        if (!load) continue;
        int m0, n0, j;
        coords(it, m0, n0, j);
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_role(&bars->empty[s], ph ^ 1u);
          tc::mbar_arrive_expect_tx(&bars->full[s], (uint32_t)stage_bytes);
          uint8_t* a = sStage + (size_t)s * stage_bytes;
          if (p.a_merge) {   // TokenMerge: k-block kb lives in quadrant (nh, nw) of the fine grid, channels e0..e0+63
            const int qd = (kb * BK) / p.mC, e0 = kb * BK - qd * p.mC;
            tc::tma_load_5d(a, &tma, &bars->full[s], e0, qd & 1, p.box_h == 1 ? m0 % p.mwc : 0, qd >> 1, m0 / p.mwc);
          } else {
            tc::tma_load_2d(a, &tma, &bars->full[s], kb * BK, m0);
          }
          if (!cfg.b_res) tc::tma_load_2d(a + A_STAGE_BYTES, &tmb, &bars->full[s], kb * BK, n0);
          if (++s == (uint32_t)cfg.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        KDB_TRACE(0);
      }
    }
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuers
    // The tensor pipe queues almost nothing behind the executing tcgen05.mma (tools/mma_dual_issue_bench.cu: a single thread that
    // spends 200 / 400 / 600 cycles between two 8-MMA tiles loses 210 / 290 / 400 cycles of tensor time per tile), and this loop's
    // bookkeeping + barrier waits are ~450 cycles per tile (profiles/r2_gemm_trace_no_waits.txt: 1000-cycle tile period for 512
    // cycles of tensor work with every wait and the whole epilogue switched off).  So TWO threads in two warps issue alternate
    // tiles: one thread's bookkeeping and mbarrier waits run while the other thread's MMAs execute (same micro-benchmark: 515
    // cycles per tile with two issuers for any of those gaps).  Tile `it` belongs to issuer it % n_iss and accumulator it % NG;
    // tcgen05.commit tracks the MMAs of the committing thread only, so a ring stage that both issuers read (weight-resident mode
    // with several n-blocks per A tile) collects one commit from each (empty[] expects two arrivals).
    // mbarrier waits test a phase PARITY: a waiter that skips phases (issuer 1 waiting for the stages of tile 1 while tile 0's are
    // still being filled; an accumulator whose uses alternate between the issuers) would pass on the wrong phase.  So the issuers
    // pass a token: issuer i starts the waits of its tile only after the other issuer has passed ALL waits of the tile before it
    // (turn[i]).  By induction every wait then sees its barrier at most one phase behind, exactly as with a single issuer, while
    // the waits themselves (~110 cycles each even when the phase is already complete) still overlap the other issuer's MMAs.
    const int id = warp - 1, n_iss = cfg.issuers;
    if (id < n_iss && tc::elect_one()) {
      if (cfg.b_res && n_local > 0) tc::mbar_wait_role(&bars->b_full, 0);
      const bool wait_acc = !(p.dbg & 8), wait_ab = !(p.dbg & (16 | 32));     // experiments (tools/gemm_probe.py): results are garbage
      const uint32_t stage_base = tc::smem_u32(sStage), b_base = tc::smem_u32(sB);
      const uint32_t n_stages = (uint32_t)cfg.stages, sbytes = (uint32_t)stage_bytes;
      const bool bres = cfg.b_res != 0;
      // ring advance of one A tile (nkb stages), carried incrementally: no division in the loop
      const uint32_t adv_s = (uint32_t)nkb % n_stages, adv_ph = ((uint32_t)nkb / n_stages) & 1u;
      // state of this issuer's first tile (it = id): divisions happen once, here
      uint32_t j = bres ? (uint32_t)id % (uint32_t)nb : 0u;                       // n-block inside the A tile
      const uint32_t pos0 = (bres ? (uint32_t)id / (uint32_t)nb : (uint32_t)id) * (uint32_t)nkb;   // ring position of the tile's A (and B) k-block 0
      uint32_t s0 = pos0 % n_stages, ph0 = (pos0 / n_stages) & 1u;
      uint32_t a0 = stage_base + s0 * sbytes;                                       // shared-memory address of ring stage s0
      uint32_t acc = (uint32_t)id % (uint32_t)NG, acc_par = (((uint32_t)id / (uint32_t)NG) & 1u) ^ 1u;
      const uint32_t adv_bytes = adv_s * sbytes, ring_bytes = n_stages * sbytes;
      const uint32_t b_step = (uint32_t)n_iss * (uint32_t)nkb * P_B_TILE_BYTES, b_wrap = (uint32_t)nb * (uint32_t)nkb * P_B_TILE_BYTES;
      uint32_t bj = b_base + j * (uint32_t)nkb * P_B_TILE_BYTES;                   // resident weights of n-block j
      auto advance_a_tile = [&]() {
        s0 += adv_s;
        a0 += adv_bytes;
        ph0 ^= adv_ph;
        if (s0 >= n_stages) {
          s0 -= n_stages;
          a0 -= ring_bytes;
          ph0 ^= 1u;
        }
      };
      uint32_t turn_par = (uint32_t)(id ^ 1);     // issuer 0's first wait passes on the fresh barrier
      for (uint32_t it = (uint32_t)id; it < (uint32_t)n_local; it += (uint32_t)n_iss) {
        KDB_TRACE(1);
        if (n_iss == 2) {
          tc::mbar_wait_role(&bars->turn[id], turn_par);
          turn_par ^= 1u;
        }
        if (wait_acc) tc::mbar_wait_role(&bars->tmem_empty[acc], acc_par);     // epilogue drained it
        tc::tc_fence_after();
        KDB_TRACE(2);
        const uint32_t d = tmem + acc * P_BN;
        // first / last tile of THIS issuer on the current A tile: it waits for the stages once and releases them once
        const bool first = j < (uint32_t)n_iss, last = j + (uint32_t)n_iss >= (uint32_t)nb;
        if (n_iss == 2 && !(first && wait_ab)) tc::mbar_arrive(&bars->turn[id ^ 1]);      // no stage waits in this tile: all waits are behind us
        uint32_t ss = s0, pp = ph0, aa = a0, bb = bres ? bj : a0 + A_STAGE_BYTES;
        // k-blocks go in pairs: both waits and all descriptor arithmetic first, then eight tcgen05.mma back to back
        for (int kb = 0; kb < nkb; kb += 2) {
          const bool two = kb + 1 < nkb;
          const uint32_t sa = ss, aa0 = aa, bb0 = bb;
          if (first && wait_ab) tc::mbar_wait_role(&bars->full[sa], pp);
          if (first && kb == 0) KDB_TRACE(13);
          aa += sbytes;
          bb += bres ? (uint32_t)P_B_TILE_BYTES : sbytes;
          if (++ss == n_stages) {
            ss = 0;
            pp ^= 1u;
            aa = stage_base;
            if (!bres) bb = stage_base + A_STAGE_BYTES;
          }
          const uint32_t sb = ss, aa1 = aa, bb1 = bb;
          if (two) {
            if (first && wait_ab) tc::mbar_wait_role(&bars->full[sb], pp);
            aa += sbytes;
            bb += bres ? (uint32_t)P_B_TILE_BYTES : sbytes;
            if (++ss == n_stages) {
              ss = 0;
              pp ^= 1u;
              aa = stage_base;
              if (!bres) bb = stage_base + A_STAGE_BYTES;
            }
          }
          if (first) tc::tc_fence_after();
          if (n_iss == 2 && first && wait_ab && kb + 2 >= nkb) tc::mbar_arrive(&bars->turn[id ^ 1]);   // the tile's last wait is behind us
          const uint64_t ad0 = tc::smem_desc_k_sw128(aa0), bd0 = tc::smem_desc_k_sw128(bb0);
          const uint64_t ad1 = tc::smem_desc_k_sw128(aa1), bd1 = tc::smem_desc_k_sw128(bb1);
          const uint32_t acc0 = (uint32_t)(kb != 0);
          tc::umma_bf16(d, ad0, bd0, IDESC, acc0);
          tc::umma_bf16(d, ad0 + 2ull, bd0 + 2ull, IDESC, 1u);
          tc::umma_bf16(d, ad0 + 4ull, bd0 + 4ull, IDESC, 1u);
          tc::umma_bf16(d, ad0 + 6ull, bd0 + 6ull, IDESC, 1u);
          if (two) {
            tc::umma_bf16(d, ad1, bd1, IDESC, 1u);
            tc::umma_bf16(d, ad1 + 2ull, bd1 + 2ull, IDESC, 1u);
            tc::umma_bf16(d, ad1 + 4ull, bd1 + 4ull, IDESC, 1u);
            tc::umma_bf16(d, ad1 + 6ull, bd1 + 6ull, IDESC, 1u);
          }
          if (last) {                                               // this issuer's last n-block on this A tile: its reads of the stages are all issued
            tc::umma_commit(&bars->empty[sa]);
            if (two) tc::umma_commit(&bars->empty[sb]);
          }
        }
        tc::umma_commit(&bars->tmem_full[acc]);
        KDB_TRACE(3);
        // this issuer's next tile: it + n_iss
        acc += (uint32_t)n_iss;
        if (acc >= (uint32_t)NG) {
          acc -= (uint32_t)NG;
          acc_par ^= 1u;
        }
        if (bres) {
          j += (uint32_t)n_iss;
          bj += b_step;
          while (j >= (uint32_t)nb) {
            j -= (uint32_t)nb;
            bj -= b_wrap;
            advance_a_tile();
          }
        } else {
          advance_a_tile();
          if (n_iss == 2) advance_a_tile();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: two ping-pong groups of 4 warps
    // group grp owns accumulator grp, staging tile grp and every tile with (it & 1) == grp, so one group's TMEM-load /
    // barrier latency overlaps the other group's arithmetic.  Thread = one accumulator row, two passes of 64 columns.
    // (Measured: 2 x 8 warps with one pass each is NOT faster -- the tile time is set by the MMA<->epilogue hand-off
    // latency, not by epilogue issue slots -- and its 96-register cap spills the QKV / residual variants.)
    const int ew = warp - 3;                 // 0 .. 4 NG - 1
    const int grp = ew >> 2;
    const int q = pwarp & 3;                 // TMEM lane quadrant this warp may touch
    const int row = q * 32 + lane;
    const bool issuer = (ew & 3) == 0 && lane == 0;
    // staging tiles: one per group, or two per group (sc_bufs == 4, residual variants when shared memory allows): the residual of
    // the group's NEXT tile is then loaded into the other buffer a whole tile period ahead instead of behind the current store
    const uint32_t nbuf = (uint32_t)cfg.sc_bufs / (uint32_t)NG;
    uint8_t* const ct_base = sC + (size_t)grp * nbuf * OUT_BYTES;
    uint8_t* ct = ct_base;
    float facv = 0.f;
    if constexpr (EPI == TCE_SPLIT) facv = __ldg(p.fac);
    // residual / skip tile of `tile` -> this group's staging tile (issuer thread only; the staging tile must be free)
    auto load_resid = [&](uint32_t it_) {
      if constexpr (RES) {
        int m0, n0, j_;
        coords(it_, m0, n0, j_);
        const uint32_t b_ = nbuf == 2 ? (it_ / (uint32_t)NG) & 1u : 0u;
        uint64_t* rf = &bars->resid_full[grp][b_];
        uint8_t* ct = ct_base + (size_t)b_ * OUT_BYTES;
        tc::mbar_arrive_expect_tx(rf, P_OUT_BYTES);
        if constexpr (EPI == TCE_SPLIT) {   // skip tensor: fine tokens of quadrant (nh, nw) = n0 / Cf, channels e0..e0+127
          const int qd = n0 / p.Cf, e0 = n0 - qd * p.Cf;
          const int wx0 = p.box_h == 1 ? m0 % p.wc : 0, bhy0 = m0 / p.wc;
          tc::tma_load_5d(ct, &tmr, rf, e0, qd & 1, wx0, qd >> 1, bhy0);
          tc::tma_load_5d(ct + SUB_TILE_BYTES, &tmr, rf, e0 + 64, qd & 1, wx0, qd >> 1, bhy0);
        } else {
          tc::tma_load_2d(ct, &tmr, rf, n0, m0);
          tc::tma_load_2d(ct + SUB_TILE_BYTES, &tmr, rf, n0 + 64, m0);
        }
      }
    };
    if constexpr (RES) {
      if (issuer && grp < n_local) load_resid((uint32_t)grp);
      if (issuer && nbuf == 2 && grp + NG < n_local) load_resid((uint32_t)(grp + NG));
    }
    // fused RMSNorm (consumer side): the row statistics of this thread's row are fetched one tile of this group AHEAD (the load
    // used to sit at the top of every tile, 15 % of all warp samples of the GEGLU kernel stalled on it: profiles/r2_ncu_*source*)
    float4 ss_pre0 = make_float4(0.f, 0.f, 0.f, 0.f), ss_pre1 = ss_pre0;
    auto prefetch_ss = [&](uint32_t it_) {
      int m0_, n0_, j2_;
      coords(it_, m0_, n0_, j2_);
      const int64_t m_ = (int64_t)m0_ + row;
      const float4* sp = reinterpret_cast<const float4*>(p.ss_in + (m_ < p.M ? m_ : 0) * SS_PARTS);
      ss_pre0 = __ldg(sp);
      ss_pre1 = __ldg(sp + 1);
    };
    if (p.ss_in != nullptr && grp < n_local) prefetch_ss((uint32_t)grp);
    uint32_t use = 0;                        // tiles this group has finished = it / NG
    for (uint32_t it = (uint32_t)grp; it < (uint32_t)n_local; it += (uint32_t)NG, ++use) {
      int m0, n0, j_;
      coords(it, m0, n0, j_);
      const uint32_t buf = nbuf == 2 ? use & 1u : 0u;
      ct = ct_base + (size_t)buf * OUT_BYTES;
      if constexpr (RES) {
        // two buffers: the store of this group's previous tile (other buffer) drained long ago -> refill it for the tile after this one
        if (nbuf == 2 && issuer && use >= 1 && it + NG < (uint32_t)n_local) {
          tc::tma_store_wait_read();
          load_resid(it + NG);
        }
      }
      const int64_t m = (int64_t)m0 + row;
      const uint32_t acc = (uint32_t)grp;
      // fused RMSNorm (consumer side): the producer of x left sum(x^2) of every token, one slot per 128 channels
      float rstd = 1.f;
      if (p.ss_in != nullptr) {
        rstd = rsqrtf(tc::rowss_sum(ss_pre0, ss_pre1, p.K >> 7) / (float)p.K + 1e-6f);
        if (it + NG < (uint32_t)n_local) prefetch_ss(it + (uint32_t)NG);
      }
      // QKV: the RoPE table row does not depend on the accumulator -> pass 0's row is fetched before waiting for the MMA, pass
      // 1's while pass 0 is being scaled / packed / stored (one live copy of the row: 32 registers instead of 64)
      float4 cs[8];
      [[maybe_unused]] auto load_cs = [&](int g) {
        const int n = n0 + g * 64;
        const int t3 = n / p.C, head = (n - t3 * p.C) >> 6;
        if (t3 < 2) {
          const int64_t tok = (m < p.M ? m : 0) % p.T;
          const float4* tb = reinterpret_cast<const float4*>(p.rope) + (int64_t)head * 8 * p.T + tok;   // [head][i][token]
#pragma unroll
          for (int i = 0; i < 8; ++i) cs[i] = __ldg(tb + (int64_t)i * p.T);
        }
      };
      if constexpr (EPI == TCE_QKV) load_cs(0);
      if (issuer) KDB_TRACE(4);
      // The staging tile is needed only at the first shared-memory store of pass 0: waiting for the previous TMA store to
      // drain is deferred until then (staging_free below), behind the accumulator wait and pass 0's arithmetic.
      // (RES: the residual load into the staging tile was issued after that drain; resid_full orders the writes.)
      auto staging_free = [&]() {
        if constexpr (!RES) {
          if (issuer) tc::tma_store_wait_read();           // this group's previous store has finished READING the staging tile
          if (issuer) KDB_TRACE(5);
          tc::named_barrier_sync(1 + 2 * grp, 128);
        }
      };
      if (issuer) KDB_TRACE(6);
      tc::mbar_wait(&bars->tmem_full[acc], use & 1u);
      tc::tc_fence_after();
      if (issuer) KDB_TRACE(7);
      float ss_acc[4] = {0.f, 0.f, 0.f, 0.f};   // producer side: sum of squares of the row this thread writes
      if constexpr (RES) tc::mbar_wait(&bars->resid_full[grp][buf], (nbuf == 2 ? use >> 1 : use) & 1u);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float v[64];
        {
          uint32_t r0[32], r1[32];
          const uint32_t taddr = tmem + acc * P_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64);
          tc::tmem_ld32_nowait(taddr, r0);
          tc::tmem_ld32_nowait(taddr + 32, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
        }
        if (g == 1) {                                      // whole accumulator row is in registers: the MMA warp may refill it
          tc::tc_fence_before();
          tc::mbar_arrive(&bars->tmem_empty[acc]);
        }
        if (issuer) KDB_TRACE(8 + g);
        if (EPI != TCE_GEGLU && p.ss_in != nullptr) {
          // fused RMSNorm row scale.  q and k are cosine-normalised afterwards (scale invariant): only v needs it.
          bool apply = true;
          if constexpr (EPI == TCE_QKV) apply = (n0 + g * 64) >= 2 * p.C;
          if (apply) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= rstd;
          }
        }
        // Epilogue arithmetic stays in fp32 and is rounded to bf16 once, at the pack (a K=128 tile leaves ~4 ALU
        // instructions per output element before the epilogue, not the tensor pipe, sets the pace).
        if constexpr (EPI == TCE_GEGLU) {
          // columns come as [8 value | 8 gate] groups (interleaved up_proj rows); packed fp32 pairs halve the issue count
          const tc::f32x2 r2 = tc::pk2(rstd, rstd), rh = tc::pk2(0.5f * rstd, 0.5f * rstd);     // the GELU's 0.5 rides on the value's row scale
          uint4 og[4];
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tc::f32x2 val = tc::pk2(v[gg * 16 + 2 * j], v[gg * 16 + 2 * j + 1]);
              tc::f32x2 gate = tc::pk2(v[gg * 16 + 8 + 2 * j], v[gg * 16 + 8 + 2 * j + 1]);
              val = tc::mul2(val, rh);
              if (p.ss_in != nullptr) gate = tc::mul2(gate, r2);
              float o0, o1;
              tc::upk2(tc::geglu2(val, gate), o0, o1);
              o[j] = tc::pack_bf16x2(o0, o1);
            }
            og[gg] = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if (g == 0) staging_free();
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) *reinterpret_cast<uint4*>(ct + tc::sw128_offset(row, g * 4 + gg)) = og[gg];
        } else {
          if constexpr (RES) {
            const uint8_t* rt = ct + g * SUB_TILE_BYTES;     // this thread reads and then overwrites only its own row
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 r4 = *reinterpret_cast<const uint4*>(rt + tc::sw128_offset(row, j));
              const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {      // bf16 -> fp32 is a shift / mask
                const float lo = __uint_as_float(rw[t] << 16), hi = __uint_as_float(rw[t] & 0xffff0000u);
                if constexpr (EPI == TCE_SPLIT) {   // torch.lerp(skip, x, fac) (reference :621)
                  const float d0 = v[j * 8 + t * 2] - lo, d1 = v[j * 8 + t * 2 + 1] - hi;
                  v[j * 8 + t * 2] = (facv < 0.5f) ? fmaf(facv, d0, lo) : v[j * 8 + t * 2] - d0 * (1.f - facv);
                  v[j * 8 + t * 2 + 1] = (facv < 0.5f) ? fmaf(facv, d1, hi) : v[j * 8 + t * 2 + 1] - d1 * (1.f - facv);
                } else {
                  v[j * 8 + t * 2] += lo;
                  v[j * 8 + t * 2 + 1] += hi;
                }
              }
            }
          }
          if constexpr (RES || EPI == TCE_STORE) {
            if (p.ss_out != nullptr) {
#pragma unroll
              for (int i = 0; i < 64; i += 4) {
                ss_acc[0] = fmaf(v[i], v[i], ss_acc[0]);
                ss_acc[1] = fmaf(v[i + 1], v[i + 1], ss_acc[1]);
                ss_acc[2] = fmaf(v[i + 2], v[i + 2], ss_acc[2]);
                ss_acc[3] = fmaf(v[i + 3], v[i + 3], ss_acc[3]);
              }
            }
          }
          if constexpr (EPI == TCE_QKV) {
            const int n = n0 + g * 64;               // one head of q, k or v (feature order (t nh e), d_head 64)
            const int t3 = n / p.C, head = (n - t3 * p.C) >> 6;
            if (t3 < 2) {
              // cosine-sim scale + axial RoPE on packed fp32 pairs.  Columns (2i, 2i+1) pair with (16+2i, 17+2i); the table
              // holds (cos_2i, cos_2i+1, sin_2i, sin_2i+1) per float4, so every operand is a natural register pair.
              tc::f32x2 P[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) P[i] = tc::pk2(v[2 * i], v[2 * i + 1]);
              tc::f32x2 q0 = tc::mul2(P[0], P[0]), q1 = tc::mul2(P[1], P[1]), q2 = tc::mul2(P[2], P[2]), q3 = tc::mul2(P[3], P[3]);
#pragma unroll
              for (int i = 4; i < 32; i += 4) {
                q0 = tc::fma2(P[i], P[i], q0);
                q1 = tc::fma2(P[i + 1], P[i + 1], q1);
                q2 = tc::fma2(P[i + 2], P[i + 2], q2);
                q3 = tc::fma2(P[i + 3], P[i + 3], q3);
              }
              float e0, e1, e2, e3, e4, e5, e6, e7;
              tc::upk2(q0, e0, e1);
              tc::upk2(q1, e2, e3);
              tc::upk2(q2, e4, e5);
              tc::upk2(q3, e6, e7);
              const float sc = sqrtf(__ldg(p.qk_scale + head)) * rsqrtf(((e0 + e1) + (e2 + e3)) + ((e4 + e5) + (e6 + e7)) + 1e-6f);
              const tc::f32x2 sc2 = tc::pk2(sc, sc);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const tc::f32x2 C = tc::pk2(cs[i].x, cs[i].y), S = tc::pk2(cs[i].z, cs[i].w);
                const tc::f32x2 NS = S ^ 0x8000000080000000ull;
                const tc::f32x2 X1 = P[i], X2 = P[8 + i];
                P[i] = tc::mul2(tc::fma2(X2, NS, tc::mul2(X1, C)), sc2);
                P[8 + i] = tc::mul2(tc::fma2(X1, S, tc::mul2(X2, C)), sc2);
              }
              if (g == 0) load_cs(1);                // the table row of pass 1 streams in behind the rest of pass 0
#pragma unroll
              for (int i = 16; i < 32; ++i) P[i] = tc::mul2(P[i], sc2);
#pragma unroll
              for (int i = 0; i < 32; ++i) tc::upk2(P[i], v[2 * i], v[2 * i + 1]);
            } else if (g == 0) {
              load_cs(1);
            }
          }
          uint8_t* cg = ct + g * SUB_TILE_BYTES;
          if (g == 0) staging_free();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(cg + tc::sw128_offset(row, j)) =
                make_uint4(tc::pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), tc::pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                           tc::pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), tc::pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
        }
      }
      if constexpr (RES || EPI == TCE_STORE) {
        if (p.ss_out != nullptr && m < p.M) {     // statistics of the NEW residual stream for the next fused RMSNorm
          const float ssv = (ss_acc[0] + ss_acc[1]) + (ss_acc[2] + ss_acc[3]);
          if constexpr (EPI == TCE_SPLIT) {       // coarse token m, quadrant qd -> fine token
            const int qd = n0 / p.Cf, e0 = n0 - qd * p.Cf;
            const int64_t bhy = m / p.wc, wx = m - bhy * p.wc;
            const int64_t fine = (2 * bhy + (qd >> 1)) * (2 * (int64_t)p.wc) + 2 * wx + (qd & 1);
            p.ss_out[fine * SS_PARTS + (e0 >> 7)] = ssv;
          } else {
            p.ss_out[m * SS_PARTS + (n0 >> 7)] = ssv;
          }
        }
      }
      if (issuer) KDB_TRACE(10);
      tc::fence_proxy_async();
      tc::named_barrier_sync(2 + 2 * grp, 128);
      if (issuer) KDB_TRACE(11);
      if (issuer) {
        if constexpr (EPI == TCE_GEGLU) {
          tc::tma_store_2d(&tmc, ct, n0 / 2, m0);
        } else if constexpr (EPI == TCE_SPLIT) {
          const int qd = n0 / p.Cf, e0 = n0 - qd * p.Cf;
          const int wx0 = p.box_h == 1 ? m0 % p.wc : 0, bhy0 = m0 / p.wc;
          tc::tma_store_5d(&tmc, ct, e0, qd & 1, wx0, qd >> 1, bhy0);
          tc::tma_store_5d(&tmc, ct + SUB_TILE_BYTES, e0 + 64, qd & 1, wx0, qd >> 1, bhy0);
        } else {
          tc::tma_store_2d(&tmc, ct, n0, m0);
          tc::tma_store_2d(&tmc, ct + SUB_TILE_BYTES, n0 + 64, m0);
        }
        tc::tma_store_commit();
        KDB_TRACE(12);
        if constexpr (RES) {   // prefetch the residual of this group's next tile as soon as the store has drained the staging tile
          if (nbuf == 1 && it + NG < (uint32_t)n_local) {
            tc::tma_store_wait_read();
            load_resid(it + NG);
          }
        }
      }
    }
    if (issuer) tc::tma_store_wait_read();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, TMEM_COLS);
  }
}

inline int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
    return v;
  }();
  return n;
}

inline size_t persist_smem(int nkb, const PersistCfg& c, int out_bytes) {
  const size_t stage = c.b_res ? A_STAGE_BYTES : A_STAGE_BYTES + P_B_TILE_BYTES;
  // (the residual tile shares the output staging tiles)
  return (c.b_res ? (size_t)c.nb * nkb * P_B_TILE_BYTES : 0) + (size_t)c.stages * stage + (size_t)c.sc_bufs * out_bytes + sizeof(PersistBars) + 1024;
}

// weight-resident when the [128 x K] block plus a >= 3-deep A ring fits; otherwise stream both operands
inline PersistCfg persist_config(int K, int n_tiles_n, bool resid, bool one_group_only, int ng, int out_bytes) {
  static const bool no_bres = [] {
    const char* e = getenv("KDB200_GEMM_NO_BRES");
    return e != nullptr && e[0] == '1';
  }();
  static const int max_nb = [] {
    const char* e = getenv("KDB200_GEMM_MAX_NB");
    return e != nullptr && e[0] >= '1' && e[0] <= '4' ? e[0] - '0' : 3;
  }();
  const int nkb = K / BK;
  if (!no_bres && n_tiles_n <= num_sms()) {
    // several resident n-blocks: the A ring must hold two whole A tiles so the next m-tile streams in behind the current one
    for (int nb = max_nb; nb >= 2; --nb) {
      if (n_tiles_n % nb != 0 || (one_group_only && nb != n_tiles_n)) continue;
      for (int st = 3 * nkb; st >= 2 * nkb; st -= nkb) {
        PersistCfg c{st, 1, ng, nb};
        if (st <= P_MAX_STAGES && persist_smem(nkb, c, out_bytes) <= P_SMEM_LIMIT) return c;
      }
    }
    if (nkb <= 6 && !(one_group_only && n_tiles_n != 1)) {
      if (resid) {   // two staging tiles per epilogue group hide the residual-load latency; worth two ring stages
        for (int st = 6; st >= 4; --st) {
          PersistCfg c{st, 1, 2 * ng, 1};
          if (st >= 2 * nkb && persist_smem(nkb, c, out_bytes) <= P_SMEM_LIMIT) return c;
        }
      }
      for (int st = 6; st >= 3; --st) {
        PersistCfg c{st, 1, ng, 1};
        if (persist_smem(nkb, c, out_bytes) <= P_SMEM_LIMIT) return c;
      }
    }
  }
  PersistCfg c{4, 0, ng, 1};
  return c;
}

template <int EPI>
int launch_persist(const bf16* A, const bf16* W, TcParams p, cudaStream_t st) {
  CUtensorMap ta, tb, tcm, tr;
  int rc;
  if (p.a_merge) {
    if ((rc = tmap_quad(&ta, A, p.mC, p.mwc, (uint64_t)p.M / p.mwc, p.box_w, p.box_h))) return rc;
  } else {
    if ((rc = tmap_2d(&ta, A, (uint64_t)p.K, (uint64_t)p.M, BK, BM))) return rc;
  }
  if ((rc = tmap_2d(&tb, W, (uint64_t)p.K, (uint64_t)p.N, BK, P_BN))) return rc;
  if (EPI == TCE_SPLIT) {
    if ((rc = tmap_quad(&tcm, p.out, p.Cf, p.wc, (uint64_t)p.M / p.wc, p.box_w, p.box_h))) return rc;
    if ((rc = tmap_quad(&tr, p.resid, p.Cf, p.wc, (uint64_t)p.M / p.wc, p.box_w, p.box_h))) return rc;
  } else {
    const uint64_t n_out = EPI == TCE_GEGLU ? (uint64_t)p.N / 2 : (uint64_t)p.N;
    if ((rc = tmap_2d(&tcm, p.out, n_out, (uint64_t)p.M, 64, BM))) return rc;
    if (EPI == TCE_RESID) {
      if ((rc = tmap_2d(&tr, p.resid, (uint64_t)p.N, (uint64_t)p.M, 64, BM))) return rc;
    } else {
      tr = ta;
    }
  }
  const int n_tiles_n = p.N / P_BN;
  // QKV: q/k tiles are heavier than v tiles, so a CTA must either own all n-blocks (resident) or take tiles in the mixed streaming order
  constexpr int NG = epi_groups<EPI>();
  constexpr int OUT_BYTES = EPI == TCE_GEGLU ? SUB_TILE_BYTES : P_OUT_BYTES;
  PersistCfg cfg = persist_config(p.K, n_tiles_n, EPI == TCE_RESID || EPI == TCE_SPLIT, EPI == TCE_QKV, NG, OUT_BYTES);
  static const int issuers = [] {
    const char* e = getenv("KDB200_GEMM_ISSUERS");
    return e != nullptr && e[0] == '1' ? 1 : 2;
  }();
  cfg.issuers = issuers;
  p.stages = cfg.stages;
  const size_t smem = persist_smem(p.K / BK, cfg, OUT_BYTES);
  static bool attr_set = false;
  if (!attr_set) {
    KDB_CUDA(cudaFuncSetAttribute(gemm_tc_persist<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_LIMIT));
    attr_set = true;
  }
  const int n_groups = cfg.b_res ? n_tiles_n / cfg.nb : n_tiles_n;
  const int64_t tiles = cfg.b_res ? (int64_t)n_groups * ceil_div(p.M, BM) : (int64_t)n_tiles_n * ceil_div(p.M, BM);
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  if (cfg.b_res) grid = grid / n_groups * n_groups;      // every CTA stays on one group of n-blocks
  static const bool trace_on = [] {
    const char* e = getenv("KDB200_GEMM_TRACE");
    return e != nullptr && e[0] == '1';
  }();
  static long long* trace_buf = nullptr;
  {
    const char* e = getenv("KDB200_GEMM_DBG");      // read per launch: tools/gemm_probe.py flips it between timed runs
    p.dbg = e != nullptr ? atoi(e) : 0;
  }
  if (trace_on) {
    if (trace_buf == nullptr) KDB_CUDA(cudaMalloc(&trace_buf, 32 * 16 * sizeof(long long)));
    KDB_CUDA(cudaMemsetAsync(trace_buf, 0, 32 * 16 * sizeof(long long), st));
    p.trace = trace_buf;
  }
  static const bool no_pdl = [] {
    const char* e = getenv("KDB200_NO_PDL");
    return e != nullptr && e[0] == '1';
  }();
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(persist_threads<EPI>());
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = no_pdl || trace_on ? 0 : 1;
  KDB_CUDA(cudaLaunchKernelEx(&lc, gemm_tc_persist<EPI>, ta, tb, tcm, tr, p, cfg));
  KDB_LAUNCH_CHECK(F_GEMM_TC, st);
  if (trace_on) {
    static long long h[32 * 16];
    KDB_CUDA(cudaMemcpyAsync(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    KDB_CUDA(cudaStreamSynchronize(st));
    const char* names[16] = {"prod_issued", "mma_pre_empty", "mma_post_empty", "mma_committed", "epi_top", "epi_store_drained", "epi_bar1",
                             "epi_tmem_full", "epi_pass0", "epi_pass1", "epi_sts_done", "epi_bar2", "epi_store_issued", "mma_full0", "mma_issued0", "mma_full1"};
    fprintf(stderr, "GEMM trace EPI=%d M=%lld N=%d K=%d grid=%d stages=%d b_res=%d nb=%d (cycles relative to tile 0 mma_pre_empty)\n", EPI, (long long)p.M,
            p.N, p.K, grid, cfg.stages, cfg.b_res, cfg.nb);
    const long long t0 = h[1];
    for (int t = 0; t < 12; ++t) {
      fprintf(stderr, " tile %2d:", t);
      for (int k = 0; k < 16; ++k) fprintf(stderr, " %s=%lld", names[k], h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}
