// Developer micro-benchmark: tensor-memory read rate with all 16 warps of a CTA reading at once (the tap-combine step of conv_tc.cu loads 6-10
// blocks of 4 columns per thread before one wait and takes ~1850 cycles: is that TMEM bandwidth or latency?).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X> __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void ld<4>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t));
}
template <> __device__ __forceinline__ void ld<16>(uint32_t t, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                 "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t));
}
// every warp issues NL loads of X columns (column stride between loads: 16), waits once; repeated REP times
template <int X, int NL>
__global__ void __launch_bounds__(512) kern(int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * X;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int rep = 0; rep < 16; ++rep) {
      uint32_t r[NL][X];
#pragma unroll
      for (int i = 0; i < NL; ++i) ld<X>(base + i * 16 * (X / 4), r[i]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < NL; ++i)
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= r[i][j];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) *out = t1 - t0;
  if (acc == 0x12345678u) *sink = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
  long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  auto run = [&](auto k, const char* name, int x, int nl) {
    for (int nw : {1, 4, 8, 16}) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) { k<<<1, 512>>>(nw, d, s); cudaDeviceSynchronize(); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost); }
      const double bytes = 16.0 * nw * 32 * nl * x * 4;
      printf("%s: %2d warps x %2d loads of %2d columns: %6.0f cycles per round, %.1f B/clk\n", name, nw, nl, x, c / 16.0, bytes / c);
    }
  };
  run(kern<4, 6>, "x4", 4, 6);
  run(kern<4, 10>, "x4", 4, 10);
  run(kern<16, 2>, "x16", 16, 2);
  run(kern<16, 4>, "x16", 16, 4);
  return 0;
}
