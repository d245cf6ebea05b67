// f32x2_bench.cu -- does packed fp32 (fma.rn.f32x2 -> FFMA2) raise the FMA rate on sm_100?  Measured on B200: 122 FMA/clk/SM either way
// (profiles/README.md).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/f32x2_bench.cu -o tools/bin/f32x2_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// mode 0: 16 independent FFMA chains x 2 (32 floats) ; mode 1: 16 independent FFMA2 chains (32 floats)
__global__ void k(int mode, int iters, float* out, long long* cyc) {
  float a[32];
  for (int i = 0; i < 32; ++i) a[i] = threadIdx.x + i;
  const float b = 1.0001f, c = 0.5f;
  long long t0 = clock64();
  if (mode == 0) {
    for (int r = 0; r < iters; ++r)
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], b, c);
  } else {
    uint64_t p[16];
    const uint64_t bb = pk(b, b), cc = pk(c, c);
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    for (int r = 0; r < iters; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) p[i] = fma2(p[i], bb, cc);
#pragma unroll
    for (int i = 0; i < 16; ++i) upk(p[i], a[2 * i], a[2 * i + 1]);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 32; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = t1 - t0;
}
int main() {
  float* o; long long* c;
  cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 16);
  for (int warps : {4, 8, 16}) for (int mode = 0; mode < 2; ++mode) {
    k<<<148, warps * 32>>>(mode, 1000, o, c);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("warps/SM=%2d mode=%s: %lld cycles for 32000 fp32 FMAs per thread -> %.2f FMA/clk/SM\n", warps, mode ? "fma.f32x2" : "fma.f32  ", h[mode], 32000.0 * warps * 32 / h[mode]);
  }
}
