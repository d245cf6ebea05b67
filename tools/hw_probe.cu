// hw_probe.cu -- two hardware facts that decide kernel designs in this repo (B200, sm_100a):
//   1. MUFU rate of ex2.approx.ftz.f32 against the packed ex2.approx.ftz.bf16x2 / f16x2 forms (is a packed exponential two results per
//      MUFU slot, i.e. does it halve the softmax's exponential time?), and tanh.approx.f32 / bf16x2 for the GEGLU epilogue;
//   2. L2 -> shared-memory bandwidth of cp.async.bulk when EVERY SM streams the SAME (L2-resident) buffer: the cost of streaming
//      weights per token tile instead of keeping them resident (fused FFN design).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/hw_probe.cu -o tools/bin/hw_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void mufu_kernel(int iters, float* out, long long* cyc) {
  // 8 independent chains per thread
  long long t0 = 0, t1 = 0;
  float s = 0.f;
  if (MODE == 0 || MODE == 3) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    t0 = clock64();
    for (int r = 0; r < iters; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0)
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        else
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
  } else {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0xbc00bc00u + threadIdx.x + i;
    t0 = clock64();
    for (int r = 0; r < iters; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 1)
          asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
        else if (MODE == 2)
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
        else
          asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
      }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float(a[i]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every CTA: `iters` passes over the same `bytes` buffer in 16 KiB bulk copies through a 4-deep ring (one thread issues, waits by mbarrier)
__global__ void __launch_bounds__(128, 1) l2_stream_kernel(const uint8_t* src, int bytes, int iters, long long* cyc, int per_cta_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CH = 16384, ST = 8;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * CH);
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const uint8_t* base = src + (size_t)per_cta_offset * blockIdx.x;
  if (threadIdx.x == 0) {
    const int nchunk = bytes / CH, total = nchunk * iters;
    long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t ph[ST] = {0};
    while (done < total) {
      while (issued < total && issued - done < ST) {
        const int s = issued % ST;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(CH));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + s * CH)),
                     "l"(base + (size_t)(issued % nchunk) * CH), "r"(CH), "r"(smem_u32(&bars[s]))
                     : "memory");
        ++issued;
      }
      const int s = done % ST;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(ph[s]));
      ph[s] ^= 1u;
      ++done;
    }
    long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
}


// TMEM read / write throughput: W warps (warp w owns lane quadrant w % 4), each loops over tcgen05.ld 32x32b.x32 (4 KiB per warp
// instruction) with two loads in flight per wait; mode 1 = tcgen05.st 32x32b.x32.
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, int mode, long long* cyc, float* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32], q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r[i] = threadIdx.x + i; q[i] = i; }
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a0 = base + (uint32_t)((it * 64) & 511), a1 = base + (uint32_t)((it * 64 + 32) & 511);
    if (mode == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(a0));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                   : "r"(a1));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[3]) + __uint_as_float(q[17]);
    } else {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
                   ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(a0));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
                   ::"r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]), "r"(q[8]), "r"(q[9]), "r"(q[10]), "r"(q[11]), "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15]), "r"(q[16]), "r"(q[17]), "r"(q[18]), "r"(q[19]), "r"(q[20]), "r"(q[21]), "r"(q[22]), "r"(q[23]), "r"(q[24]), "r"(q[25]), "r"(q[26]), "r"(q[27]), "r"(q[28]), "r"(q[29]), "r"(q[30]), "r"(q[31]), "r"(a1));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_slot));
}

int main() {
  float* o;
  long long* c;
  cudaMalloc(&o, 148 * 1024 * 4);
  cudaMalloc(&c, 148 * 8);
  const char* names[5] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "ex2.approx.f16x2", "tanh.approx.f32", "tanh.approx.bf16x2"};
  for (int warps : {4, 8, 16})
    for (int mode = 0; mode < 5; ++mode) {
      const int iters = 2000;
      switch (mode) {
        case 0: mufu_kernel<0><<<148, warps * 32>>>(iters, o, c); break;
        case 1: mufu_kernel<1><<<148, warps * 32>>>(iters, o, c); break;
        case 2: mufu_kernel<2><<<148, warps * 32>>>(iters, o, c); break;
        case 3: mufu_kernel<3><<<148, warps * 32>>>(iters, o, c); break;
        default: mufu_kernel<4><<<148, warps * 32>>>(iters, o, c); break;
      }
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double instr = 8.0 * iters * warps * 32;
      printf("MUFU warps/SM=%2d %-22s %8lld cycles -> %.2f thread-instructions/clk/SM (%s results/clk/SM: %.2f)\n", warps, names[mode], h, instr / h,
             (mode == 1 || mode == 2 || mode == 4) ? "2 per instr," : "1 per instr,", instr / h * ((mode == 1 || mode == 2 || mode == 4) ? 2 : 1));
    }
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 12, 16}) {
      const int iters = 4000;
      tmem_kernel<<<148, warps * 32>>>(iters, mode, c, o);
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("TMEM %s warps/SM=%2d: %8lld cycles -> %.1f B/clk/SM  (a 128x128 fp32 tile = 64 KiB takes %.0f cycles)  %s\n", mode ? "tcgen05.st" : "tcgen05.ld", warps, h,
             8192.0 * iters * warps / h, 65536.0 / (8192.0 * iters * warps / h), cudaGetErrorString(cudaGetLastError()));
    }
  // L2 streaming
  uint8_t* buf;
  const size_t big = (size_t)148 * 512 * 1024;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int shared_buf = 1; shared_buf >= 0; --shared_buf)
    for (int kb : {128, 288, 512}) {
      const int iters = 200;
      l2_stream_kernel<<<148, 128, 8 * 16384 + 256>>>(buf, kb * 1024, 5, c, shared_buf ? 0 : 512 * 1024);      // warm L2
      cudaEventRecord(e0);
      l2_stream_kernel<<<148, 128, 8 * 16384 + 256>>>(buf, kb * 1024, iters, c, shared_buf ? 0 : 512 * 1024);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long h[148];
      cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)kb * 1024 * iters;
      printf("L2->smem bulk copy, %s %3d KiB buffer: %.1f B/clk/SM (slowest CTA), chip %.2f TB/s (%.3f ms)\n",
             shared_buf ? "ALL SMs read the SAME" : "each SM its OWN    ", kb, bytes / mx, bytes * 148 / (ms * 1e-3) / 1e12, ms);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
