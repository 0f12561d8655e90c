// Developer micro-benchmark: how long does an L2-resident 80 KB weight image take to reach shared memory?
// (chain64.cu waits ~6-8 k cycles for it.)  Variants: cp.async.bulk in n chunks issued by one thread, cp.async.bulk issued by n different
// threads (one chunk each), plain 16-byte loads by all 512 threads; 1 / 32 / 128 CTAs reading the SAME image or one image each.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) kern(const uint8_t* src, size_t cta_stride, int bytes, int nchunk, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const uint8_t* s = src + (size_t)blockIdx.x * cta_stride;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  const int per = bytes / nchunk;
  if (mode == 0) {
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      for (int i = 0; i < nchunk; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + i * per)), "l"(s + i * per), "r"(per),
                     "r"(smem_u32(&bar)) : "memory");
    }
  } else if (mode == 1) {
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    __syncthreads();
    if ((tid & 31) == 0 && (tid >> 5) < nchunk) {
      const int i = tid >> 5;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + i * per)), "l"(s + i * per), "r"(per),
                   "r"(smem_u32(&bar)) : "memory");
    }
  } else if (mode == 3 || mode == 4 || mode == 5) {
    // interference test: thread 0 copies while warps 1..15 keep writing a scratch area of shared memory (mode 3: + fence.proxy.async after
    // every store, as an epilogue that prepares the next A operand does; mode 4: stores only; mode 5: fences only)
    __shared__ uint4 scratch[512];
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(s), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    if (tid >= 32) {
      for (int it = 0; it < nchunk; ++it) {
        if (mode != 5) scratch[tid] = make_uint4(it, tid, 0, 0);
        if (mode != 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    }
  } else {
    for (int i = tid; i < bytes / 16; i += 512) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(s) + i);
    __syncthreads();
  }
  if (mode != 2) asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(&bar)) : "memory");
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  const int bytes = 80 * 1024;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, (size_t)bytes * 148);
  cudaMemset(src, 1, (size_t)bytes * 148);
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024);
  {  // cold L2: a 256 MiB write between launches evicts the image -> the copy comes from DRAM
    uint8_t* flush; cudaMalloc(&flush, 256u << 20);
    for (int grid : {1, 32}) for (int nchunk : {1, 10}) {
      long long h[148]; long long sum = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(flush, rep, 256u << 20);
        kern<<<grid, 512, bytes + 1024>>>(src, 0, bytes, nchunk, 0, out);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      for (int i = 0; i < grid; ++i) sum += h[i];
      printf("COLD L2: CTAs=%3d same image cp.async.bulk chunks=%2d: mean %6lld cycles\n", grid, nchunk, sum / grid);
    }
    cudaFree(flush);
  }
  for (int mode : {3, 4, 5})
    for (int iters : {1, 8, 64}) {
      long long h[148];
      for (int rep = 0; rep < 3; ++rep) { kern<<<32, 512, bytes + 1024>>>(src, 0, bytes, iters, mode, out); cudaDeviceSynchronize(); }
      cudaMemcpy(h, out, 32 * 8, cudaMemcpyDeviceToHost);
      long long sum = 0; for (int i = 0; i < 32; ++i) sum += h[i];
      printf("INTERFERENCE (32 CTAs, hot L2): %s x %2d per thread while the copy is in flight: mean %6lld cycles\n", mode == 3 ? "st.shared + fence.proxy.async" : (mode == 4 ? "st.shared only" : "fence.proxy.async only"), iters, sum / 32);
    }
  for (int grid : {1, 32, 128})
    for (int same : {1, 0})
      for (int mode : {0, 1, 2})
        for (int nchunk : {1, 2, 10, 16}) {
          if (mode == 2 && nchunk != 1) continue;
          if (bytes % (nchunk * 16)) continue;
          long long h[148];
          for (int rep = 0; rep < 3; ++rep) {
            kern<<<grid, 512, bytes + 1024>>>(src, same ? 0 : bytes, bytes, nchunk, mode, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          }
          cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0, sum = 0;
          for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
          printf("CTAs=%3d %s image  %-34s chunks=%2d: mean %6lld max %6lld cycles (%.1f B/clk per CTA)\n", grid, same ? "same" : "own ",
                 mode == 0 ? "cp.async.bulk, one thread" : (mode == 1 ? "cp.async.bulk, one thread per chunk" : "ld.global.v4 by 512 threads"), nchunk, sum / grid, mx,
                 (double)bytes / (sum / grid));
        }
  return 0;
}
