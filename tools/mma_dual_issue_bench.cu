// mma_dual_issue_bench.cu -- the tensor pipe accepts only ~1-2 queued tcgen05.mma, so every serial gap in the issuing thread's instruction
// stream (per-tile bookkeeping: ~400-500 cycles in the persistent GEMM, profiles/r2_gemm_trace_no_waits.txt) is a bubble.  Two remedies:
//   (a) TWO issuing threads (different warps) that take alternate tiles: one thread's bookkeeping overlaps the other's MMAs;
//   (b) N = 256 instructions: twice the tensor work per issue and per bookkeeping round.
// Each tile = 8 MMAs (two 64-wide k-blocks); `chain` dependent integer multiply-adds per tile emulate the bookkeeping.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I k-diffusion_b200/csrc -I include tools/mma_dual_issue_bench.cu -o tools/bin/mma_dual_issue_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace kdb;

template <int N>
__global__ void __launch_bounds__(128, 1) dual(int tiles, int issuers, int chain, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done[2], tile_bar[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  constexpr int BT = N * 128;      // bytes of one B k-block tile
  for (int i = threadIdx.x; i < (4 * 16384 + 2 * BT) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&done[i], 1);
      tc::mbar_init(&tile_bar[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  constexpr uint32_t IDESC = tc::idesc_bf16(128, N);
  // issuer 0 = warp 3, issuer 1 = warp 2 (different schedulers)
  const int id = warp == 3 ? 0 : (warp == 2 ? 1 : -1);
  if (id >= 0 && id < issuers && tc::elect_one()) {
    const uint32_t a_base = tc::smem_u32(base), b_base = a_base + 4 * 16384;
    uint32_t as = (uint32_t)id, x = 12345u + (uint32_t)id;
    const long long t0 = clock64();
    for (int t = id; t < tiles; t += issuers) {
      // bookkeeping stand-in: `chain` dependent IMADs; the result feeds the descriptor so it cannot be hoisted or dropped
      for (int c = 0; c < chain; ++c) x = x * 1664525u + 1013904223u;
      as = (as + (x & 1u) * 0u + 1u) & 3u;
      const uint32_t d = tmem + (uint32_t)((t & 1) * (N <= 128 ? 128 : 256));
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint32_t aa = a_base + ((as + kb) & 3u) * 16384u, bb = b_base + (uint32_t)(kb * BT);
        const uint64_t ad = tc::smem_desc_k_sw128(aa), bd = tc::smem_desc_k_sw128(bb);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(d, ad + 2ull * k, bd + 2ull * k, IDESC, (uint32_t)((kb | k) != 0));
      }
      tc::umma_commit(&tile_bar[id]);
    }
    tc::umma_commit(&done[id]);
    tc::mbar_wait(&done[id], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[id] = t1 - t0;
    if (x == 42u) out[7] = 1;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run(int issuers, int chain, long long* dout) {
  const size_t smem = 4 * 16384 + 2 * N * 128 + 1024;
  cudaFuncSetAttribute(dual<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 512;
  cudaMemset(dout, 0, 64);
  dual<N><<<148, 128, smem>>>(tiles, issuers, chain, dout);
  cudaError_t le = cudaGetLastError(), e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
  const long long T = h[0] > h[1] ? h[0] : h[1];
  printf("N=%3d issuers=%d bookkeeping chain=%3d: %7.1f cycles per tile (tensor work %d) -> %5.0f FLOP/clk/SM  [%s %s]\n", N, issuers, chain, (double)T / tiles, 8 * (N / 2),
         2.0 * 128 * N * 16 * 8.0 * tiles / T, cudaGetErrorString(le), cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  for (int chain : {0, 50, 100, 150})
    for (int issuers : {1, 2}) {
      run<128>(issuers, chain, dout);
      run<256>(issuers, chain, dout);
    }
  return 0;
}
