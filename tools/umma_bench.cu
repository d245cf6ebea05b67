// umma_bench.cu -- how long does one tcgen05.mma (cta_group::1, kind::f16, M=128) take for N = 64 / 128 / 256 when the
// operands are already in shared memory?  One CTA per SM, one thread issues `reps` x 4 MMAs (one 64-wide k-block), commits
// and waits; cycles per MMA = (t1 - t0) / (4 * reps).  Also measured with a concurrent warp streaming shared memory
// (ld.shared.v4) to see how much the tensor pipe depends on free shared-memory bandwidth.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I k-diffusion_b200/csrc tools/umma_bench.cu -o tools/bin/umma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace kdb;

template <int N>
__global__ void __launch_bounds__(320, 1) umma_rate(int reps, int smem_noise, int kind, long long* out, const uint8_t* gsrc) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                 // [128 x 64] bf16, SW128
  uint8_t* sB = base + 16384;         // [N x 64]
  uint8_t* sN = base + 16384 + 32768; // 64 KB scratch the noise warps read
  __shared__ uint64_t bar;
  __shared__ uint64_t nbar[8];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) tc::mbar_init(&nbar[i], 1);
    tc::fence_barrier_init();
    stop = 0;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  constexpr uint32_t IDESC = tc::idesc_bf16(128, N);
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint64_t adesc = tc::smem_desc_k_sw128(tc::smem_u32(sA));
      const uint64_t bdesc = tc::smem_desc_k_sw128(tc::smem_u32(sB));
      // warm-up
      for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem, adesc + 2ull * k, bdesc + 2ull * k, IDESC, 1u);
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + (uint32_t)((r & 1) * (N <= 256 ? 256 : 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(d, adesc + 2ull * k, bdesc + 2ull * k, IDESC, 1u);
      }
      const long long t_issue = clock64();
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, 1);
      const long long t1 = clock64();
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t_issue - t0;
      }
      stop = 1;
    }
  } else if (warp >= 2 && warp < 2 + smem_noise && kind == 1) {
    // TMEM read noise: tcgen05.ld 32x32b.x32 of columns 384..511 (never written by the MMAs), lane quadrant warp % 4
    float v[32];
    float acc = 0.f;
    long long n = 0;
    const long long t0 = clock64();
    while (!stop) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384u + (uint32_t)(c * 32), v);
        acc += v[0] + v[31];
      }
      n += 4;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 64) {
      out[2] = n;
      out[3] = t1 - t0;
    }
    if (acc == 1.2345f) out[7] = 1;
  } else if (warp >= 2 && warp < 2 + smem_noise && kind == 2) {
    // shared-memory write noise: swizzled 16-byte stores like the epilogue's staging writes
    int off = (threadIdx.x & 31) * 16;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(tc::smem_u32(sN + ((off + i * 512) & 65535))), "r"(i) : "memory");
      off = (off + 8192) & 65535;
    }
  } else if (warp >= 2 && warp < 2 + smem_noise && kind == 3) {
    // FP32 / MUFU noise
    float a = (float)threadIdx.x, b = 1.0001f, c = 0.f;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        a = fmaf(a, b, c);
        c = fmaf(c, b, a);
        if ((i & 7) == 0) b = tc::gelu_fast(b) + 1.f;
      }
    }
    if (a + c == 1.2345f) out[7] = 1;
  } else if (warp >= 2 && warp < 2 + smem_noise && kind == 4) {
    // FP32 noise with instruction-level parallelism (16 independent chains): can issue every cycle like the GEMM epilogue
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i);
    const float b = 1.0001f, c = 0.5f;
    while (!stop) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += a[i];
    if (t == 1.2345f) out[7] = 1;
  } else if (warp >= 2 && warp < 2 + smem_noise && kind == 5) {
    // bulk-copy noise: one thread per warp keeps a 16 KiB global -> shared bulk copy in flight (async-proxy writes, like TMA loads)
    if ((threadIdx.x & 31) == 0) {
      uint64_t* nb = &nbar[warp - 2];
      uint8_t* dst = sN + ((warp - 2) & 3) * 16384;
      const uint8_t* src = gsrc + ((size_t)blockIdx.x * 8 + (warp - 2)) * 16384;
      uint32_t ph = 0;
      while (!stop) {
        tc::mbar_arrive_expect_tx(nb, 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src),
                     "r"(16384), "r"(tc::smem_u32(nb))
                     : "memory");
        tc::mbar_wait(nb, ph);
        ph ^= 1u;
      }
    }
  } else if (warp >= 2 && warp < 2 + smem_noise) {
    // shared-memory read noise: conflict-free 16-byte loads, as fast as the warp can issue them
    uint32_t acc = 0;
    int off = (threadIdx.x & 31) * 16;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tc::smem_u32(sN + ((off + i * 512) & 65535))));
        acc += v0 ^ v3 ^ v1 ^ v2;
      }
      off = (off + 8192) & 65535;
    }
    if (acc == 0x12345678u) out[7] = acc;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run(int reps, int noise, int kind, long long* dout) {
  const size_t smem = 16384 + 32768 + 65536 + 1024;
  cudaFuncSetAttribute(umma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  static const char* kinds[6] = {"ld.shared", "tcgen05.ld", "st.shared", "fma+mufu", "fma-ilp16", "bulk-copy"};
  static uint8_t* gsrc = nullptr;
  if (gsrc == nullptr) cudaMalloc(&gsrc, (size_t)160 * 8 * 16384);
  for (int grid : {sms}) {
    cudaMemset(dout, 0, 64);
    umma_rate<N><<<grid, 320, smem>>>(reps, noise, kind, dout, gsrc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4] = {0, 0, 0, 0};
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (4.0 * reps);
    printf("N=%3d grid=%3d noise=%-10s x%d warps: %7.1f cycles per 128xNx16 MMA -> %6.0f FLOP/clk/SM", N, grid, kinds[kind], noise, per, 2.0 * 128 * N * 16 / per);
    if (kind == 1 && h[3] > 0) printf("   (one noise warp: %.1f cycles per tcgen05.ld.32x32b.x32 = 4 KiB)", (double)h[3] / (double)h[2]);
    printf("  [%s]\n", cudaGetErrorString(e));
  }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  run<128>(2048, 0, 0, dout);
  run<256>(2048, 0, 0, dout);
  for (int kind = 3; kind < 6; ++kind)
    for (int noise : {4, 8}) {
      run<128>(2048, noise, kind, dout);
      run<256>(2048, noise, kind, dout);
    }
  return 0;
}
