// Developer micro-benchmark: do the packed fp32 instructions (FADD2 / FFMA2, PTX add/fma.rn.f32x2) free issue slots?
// 16 warps per CTA (4 per scheduler), one CTA per SM; per thread 8 independent accumulator pairs; variants:
//   0: 2N FADD    1: N FADD2    2: 2N FADD + N IMAD    3: N FADD2 + N IMAD   (N = 4096 per thread)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void __launch_bounds__(512) kern(float* out, long long* cyc, float inc, int ii) {
  float a[16]; unsigned long long p[8]; int q[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); q[i] = threadIdx.x + i; }
  const unsigned long long pinc = pk(inc, inc);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 512; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(inc));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], pinc);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(q[i]) : "r"(ii));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y + q[i]; }
  out[blockIdx.x * 512 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[4] = {"8192 FADD", "4096 FADD2", "8192 FADD + 4096 IMAD", "4096 FADD2 + 4096 IMAD"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) kern<0><<<148, 512>>>(out, cyc, 1.f, 3); if (m == 1) kern<1><<<148, 512>>>(out, cyc, 1.f, 3);
      if (m == 2) kern<2><<<148, 512>>>(out, cyc, 1.f, 3); if (m == 3) kern<3><<<148, 512>>>(out, cyc, 1.f, 3);
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-26s per thread, 4 warps per scheduler: %lld cycles\n", names[m], h);
  }
  return 0;
}
