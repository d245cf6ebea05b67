// mma_issue_bench.cu -- why does the persistent GEMM issue one 128x128x16 tcgen05.mma per ~100 cycles when the same instruction runs at
// 64 cycles in tools/umma_bench.cu?  One CTA per SM, one elected thread issues `tiles` x 8 MMAs (two 64-wide k-blocks per tile, like K=128)
// with NO data dependencies at all (operands are whatever is in shared memory), in four variants:
//   vary=0  constant descriptors (umma_bench's loop)
//   vary=1  A descriptor walks a ring of 8 tiles, B walks 4 resident tiles: addresses computed at run time like the GEMM's loop
//   commit=1 additionally tcgen05.commit to an mbarrier after every tile (as the GEMM does for tmem_full)
// for N = 128 and N = 256.  Reports cycles per MMA (clock64 of the issuing thread, issue start -> all MMAs complete).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I k-diffusion_b200/csrc tools/mma_issue_bench.cu -o tools/bin/mma_issue_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace kdb;

template <int N>
__global__ void __launch_bounds__(384, 1) issue_rate(int tiles, int vary, int commit, int a_tmem, long long* out, int noise, uint8_t* gdst) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  // A ring: 8 x 16 KiB at 0; B: 4 x (N x 128 B) after it
  __shared__ uint64_t bar, tile_bar, never_bar;
  __shared__ volatile int stop;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (8 * 16384 + (N <= 128 ? 4 : 2) * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init(&tile_bar, 1);
    tc::mbar_init(&never_bar, 1);
    stop = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  constexpr uint32_t IDESC = tc::idesc_bf16(128, N);
  if (warp == 0 && tc::elect_one()) {
    const uint32_t a_base = tc::smem_u32(base), b_base = a_base + 8 * 16384;
    uint32_t as = 0, bs = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (uint32_t)((t & 1) * (N <= 128 ? 128 : 256));
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t aa = a_base, bb = b_base;
        if (vary) {
          aa = a_base + as * 16384u;
          bb = b_base + bs * (uint32_t)(N * 128);
          as = (as + 1) & 7u;
          bs = (bs + 1) & (N <= 128 ? 3u : 1u);
        }
        const uint64_t ad = tc::smem_desc_k_sw128(aa), bd = tc::smem_desc_k_sw128(bb);
        if (a_tmem) {        // A from tensor memory (columns 384..447: never written, timing only)
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_bf16_ts(d, tmem + 384u + (uint32_t)(kb * 32 + k * 8), bd + 2ull * k, IDESC, (uint32_t)((kb | k) != 0));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_bf16(d, ad + 2ull * k, bd + 2ull * k, IDESC, (uint32_t)((kb | k) != 0));
        }
      }
      if (commit) tc::umma_commit(&tile_bar);
    }
    const long long t_issue = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t_issue - t0;
    }
    stop = 1;
  } else if (warp >= 4 && noise != 0) {
    // eight noise warps (warps 4..11), each iteration followed by ~100 cycles of idle time unless noted
    uint8_t* scratch = base + 8 * 16384 + (N <= 128 ? 4 : 2) * N * 128 + (warp - 4) * 2048;      // 2 KiB per warp
    long long n = 0;
    while (!stop) {
      switch (noise) {
        case 1: tc::fence_proxy_async(); break;                                                   // generic -> async proxy fence
        case 2: tc::mbar_try_wait(&never_bar, 0); break;                                           // parked mbarrier wait (suspend hint)
        case 3: tc::tc_fence_before(); tc::tc_fence_after(); break;
        case 4: __syncwarp(); break;                                                                // (bar.sync would deadlock at the stop flag)
        case 5: {                                                                                  // st.shared + proxy fence + bulk store + wait (the epilogue's tail)
          *reinterpret_cast<uint4*>(scratch + (threadIdx.x & 31) * 16) = make_uint4(1, 2, 3, 4);
          tc::fence_proxy_async();
          __syncwarp();
          if ((threadIdx.x & 31) == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(gdst + ((size_t)blockIdx.x * 8 + (warp - 4)) * 512), "r"(tc::smem_u32(scratch)) : "memory");
            tc::tma_store_commit();
            tc::tma_store_wait_read();
          }
          __syncwarp();
        } break;
        case 6: {                                                                                  // tcgen05.ld of columns the MMAs never write
          uint32_t r[32];
          tc::tmem_ld32_nowait(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 448u, r);
          tc::tmem_ld_wait(r);
          if (r[0] == 0x12345u) out[7] = 1;
        } break;
        case 7: {                                                                                  // swizzled 16-byte shared stores, back to back
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(scratch + ((threadIdx.x & 31) * 16 + i * 512) % 2048) = make_uint4(i, 2, 3, 4);
        } break;
        default: break;
      }
      if (noise != 7) __nanosleep(40);
      ++n;
    }
    if (n == 0x7fffffffffffffffll) out[6] = n;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run(int vary, int commit, int a_tmem, long long* dout, int noise = 0) {
  static uint8_t* gdst = nullptr;
  if (gdst == nullptr) cudaMalloc(&gdst, 148 * 8 * 512);
  const size_t smem = 8 * 16384 + (N <= 128 ? 4 : 2) * N * 128 + 8 * 2048 + 1024;
  if (smem > 227 * 1024) { printf("N=%d: needs %zu bytes of shared memory, skipped\n", N, smem); return; }
  cudaFuncSetAttribute(issue_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 512;
  cudaMemset(dout, 0, 64);
  issue_rate<N><<<148, 384, smem>>>(tiles, vary, commit, a_tmem, dout, noise, gdst);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(le)); return; }
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  static const char* nn[8] = {"none", "fence.proxy.async", "mbarrier.try_wait (parked)", "tcgen05.fence", "bar.sync 128", "st.shared+fence+bulk store+wait", "tcgen05.ld", "st.shared.v4 stream"};
  printf("noise=%-32s N=%3d A=%s vary=%d commit=%d: %6.1f cycles per MMA (issue loop alone %6.1f) -> %5.0f FLOP/clk/SM  [%s]\n",
         nn[noise], N, a_tmem ? "tmem" : "smem", vary, commit, (double)h[0] / (8.0 * tiles), (double)h[1] / (8.0 * tiles), 2.0 * 128 * N * 16 * 8.0 * tiles / h[0], cudaGetErrorString(e));
  (void)nn;
  fflush(stdout);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem) {
    run<128>(1, 1, a_tmem, dout);
    run<256>(1, 1, a_tmem, dout);
  }
  for (int noise = 1; noise < 8; ++noise) {
    run<128>(1, 1, 0, dout, noise);
    run<128>(1, 1, 1, dout, noise);
  }
  return 0;
}
