// mma_proto_bench.cu -- the persistent GEMM's MMA <-> epilogue hand-shake in isolation (no TMA, no data): does the PROTOCOL cost the
// ~1000 cycles per 128x128 (K=128) tile that the real kernel shows with every epilogue action switched off (profiles/r2_gemm_trace_no_waits.txt)?
//   warp 0 (elected thread): per tile  [wait tmem_empty[acc]] -> tcgen05.fence::after -> 8 x tcgen05.mma (N=128) -> commit tmem_full[acc]
//   NG groups of 4 warps:    per tile  wait tmem_full[g] -> fence::after -> [tcgen05.ld 128 columns] -> fence::before -> arrive tmem_empty[g]
//                                      -> [fence.proxy.async + bar.sync 128]  (the staging hand-off of the real epilogue)
// Variants are bit flags of `mode`: 1 = MMA waits for tmem_empty, 2 = epilogue loads the accumulator, 4 = epilogue does the proxy fence +
// named barrier, 8 = MMA thread executes fence::after every tile, 16 = MMA warp is the LAST warp of the CTA instead of warp 0.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I k-diffusion_b200/csrc -I include tools/mma_proto_bench.cu -o tools/bin/mma_proto_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace kdb;

template <int NG>
__global__ void __launch_bounds__(32 + 128 * NG, 1) proto(int tiles, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t tmem_full[4], tmem_empty[4], done;
  __shared__ uint32_t tmem_base;
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = 1 + 4 * NG;
  const int warp = (mode & 16) ? (pwarp == nwarps - 1 ? 0 : pwarp + 1) : pwarp;      // role index: 0 = MMA, 1.. = epilogue
  for (int i = threadIdx.x; i < (8 * 16384 + 4 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int a = 0; a < NG; ++a) {
      tc::mbar_init(&tmem_full[a], 1);
      tc::mbar_init(&tmem_empty[a], 128);
    }
    tc::mbar_init(&done, 1);
    tc::fence_barrier_init();
  }
  if (pwarp == 0) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  constexpr uint32_t IDESC = tc::idesc_bf16(128, 128);
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t a_base = tc::smem_u32(base), b_base = a_base + 8 * 16384;
      uint32_t as = 0, bs = 0, acc = 0, par = 1;
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t) {
        if (mode & 1) tc::mbar_wait(&tmem_empty[acc], par);
        if (mode & 8) tc::tc_fence_after();
        const uint32_t d = tmem + acc * 128u;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t aa = a_base + as * 16384u, bb = b_base + bs * 16384u;
          as = (as + 1) & 7u;
          bs = (bs + 1) & 3u;
          const uint64_t ad = tc::smem_desc_k_sw128(aa), bd = tc::smem_desc_k_sw128(bb);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_bf16(d, ad + 2ull * k, bd + 2ull * k, IDESC, (uint32_t)((kb | k) != 0));
        }
        tc::umma_commit(&tmem_full[acc]);
        if (++acc == (uint32_t)NG) {
          acc = 0;
          par ^= 1u;
        }
      }
      tc::umma_commit(&done);
      tc::mbar_wait(&done, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (mode & 1) {      // (without the hand-shake the groups would lose phases: they sit out)
    const int ew = warp - 1, grp = ew >> 2, q = pwarp & 3;
    uint32_t use = 0;
    float accv = 0.f;
    for (int t = grp; t < tiles; t += NG, ++use) {
      tc::mbar_wait(&tmem_full[grp], use & 1u);
      tc::tc_fence_after();
      if (mode & 2) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t r0[32], r1[32];
          const uint32_t taddr = tmem + (uint32_t)grp * 128u + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64);
          tc::tmem_ld32_nowait(taddr, r0);
          tc::tmem_ld32_nowait(taddr + 32, r1);
          tc::tmem_ld_wait();
          accv += __uint_as_float(r0[3]) + __uint_as_float(r1[5]);
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tmem_empty[grp]);
      if (mode & 4) {
        tc::fence_proxy_async();
        tc::named_barrier_sync(1 + grp, 128);
      }
    }
    if (accv == 1.2345f) out[7] = lane;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (pwarp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int NG>
void run(int mode, long long* dout) {
  const size_t smem = 12 * 16384 + 1024;
  cudaFuncSetAttribute(proto<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 600;
  cudaMemset(dout, 0, 64);
  proto<NG><<<148, 32 + 128 * NG, smem>>>(tiles, mode, dout);
  cudaError_t le = cudaGetLastError(), e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
  printf("groups=%d mode=%2d [%s%s%s%s%s]: %7.1f cycles per tile (8 MMAs = 512 cycles of tensor work)  [%s %s]\n", NG, mode, mode & 1 ? "wait-empty " : "", mode & 2 ? "tmem-ld " : "",
         mode & 4 ? "proxy-fence+bar " : "", mode & 8 ? "mma-fence " : "", mode & 16 ? "mma-last-warp" : "", (double)h / tiles, cudaGetErrorString(le), cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  for (int mode : {0, 8, 1, 9, 3, 11, 7, 15, 31, 27, 23}) {
    run<2>(mode, dout);
    run<3>(mode, dout);
  }
  return 0;
}
