// Developer micro-benchmark: what paces back-to-back tcgen05.mma instructions of small N?
// scripts/ts_mma_test.cu measured ~105 cycles per M = 128, K = 16 instruction for every N <= 192 with ONE accumulator chain.  This
// program separates the candidates: dependent accumulation (nacc independent accumulators in the warp-uniform issue loop), the
// M = 64 form, two issuing warps, and the same instruction stream fully unrolled.  Timing only (operands are whatever shared memory holds).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_mn(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
// V2: the leader predicate is computed once outside the loop and passed in (no elect.sync per instruction)
__device__ __forceinline__ void mma_lead(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc, uint32_t leader) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 e, %5, 0;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void mma_plain(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t e;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(e));
  return e;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(bar)), "r"(ph) : "memory");
}

// mode 0: one warp issues `total` MMAs round-robin over nacc accumulators (each N columns wide)
// mode 1: warps 0 and 1 each issue total/2 MMAs into their own accumulator
__global__ void __launch_bounds__(128) kern(int M, int N, int nacc, int total, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // 4 chunks x 16 KB
  uint8_t* sB = smem + 4 * 16384;   // 4 chunks x 32 KB
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  const uint32_t id = idesc_mn(M, N);
  long long t0 = clock64();
  if (mode == 0 && warp == 0) {
    int acc_i = 0;
    for (int i = 0; i < total; ++i) {
      const int ch = (i >> 2) & 3, k = i & 3;
      const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)) + 2 * k, bd = umma_desc(smem_u32(sB + ch * 32768)) + 2 * k;
      mma_elect(tb + acc_i * N, ad, bd, id, i >= nacc ? 1u : 0u);
      acc_i = acc_i + 1 == nacc ? 0 : acc_i + 1;
    }
    if (lane == 0) commit(&bar[0]);
    __syncwarp();
  }
  // modes 2..5: the nested loop structure of the production kernels (chunks outer, 4 K steps unrolled inside)
  if (mode >= 2 && warp == 0) {
    const uint32_t leader = elect_one();
    const int chunks = total / 4;
    if (mode == 4) {
      if (leader) {
        for (int c = 0; c < chunks; ++c) {
          const int ch = c & 3;
          const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)), bd = umma_desc(smem_u32(sB + ch * 32768));
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_plain(tb, ad + 2 * k, bd + 2 * k, id, (c == 0 && k == 0) ? 0u : 1u);
        }
        commit(&bar[0]);
      }
      __syncwarp();
    } else {
      for (int c = 0; c < chunks; ++c) {
        const int ch = c & 3;
        const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)), bd = umma_desc(smem_u32(sB + ch * 32768));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (c == 0 && k == 0) ? 0u : 1u;
          if (mode == 2) mma_elect(tb, ad + 2 * k, bd + 2 * k, id, acc);
          else mma_lead(tb, ad + 2 * k, bd + 2 * k, id, acc, leader);
        }
      }
      if (leader) commit(&bar[0]);
      __syncwarp();
    }
  }
  if (mode == 1 && warp < 2) {
    for (int i = 0; i < total / 2; ++i) {
      const int ch = (i >> 2) & 3, k = i & 3;
      const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)) + 2 * k, bd = umma_desc(smem_u32(sB + ch * 32768)) + 2 * k;
      mma_elect(tb + warp * 256, ad, bd, id, i >= 1 ? 1u : 0u);
    }
    if (lane == 0) commit(&bar[warp]);
    __syncwarp();
  }
  wait_bar(&bar[0], 0);
  if (mode == 1) wait_bar(&bar[1], 0);
  long long t1 = clock64();
  if (tid == 0) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main() {
  long long* dC;
  cudaMalloc(&dC, 8);
  const int smem = 4 * 16384 + 4 * 32768 + 2048;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int total = 512;
  auto run = [&](int M, int N, int nacc, int mode) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) {
      kern<<<1, 128, smem>>>(M, N, nacc, total, mode, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("M=%d N=%d nacc=%d mode=%d: %s\n", M, N, nacc, mode, cudaGetErrorString(e)); exit(1); }
      cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    }
    printf("M=%3d N=%3d %s=%d: %.1f cycles per MMA\n", M, N, mode ? "issuing warps" : "accumulators", mode ? 2 : nacc, (double)c / total);
  };
  for (int N : {16, 48, 96, 128, 192, 256})
    for (int nacc : {1, 2, 4})
      if (nacc * N <= 512) run(128, N, nacc, 0);
  for (int N : {48, 96, 192, 256})
    for (int nacc : {1, 2})
      if (nacc * N <= 512) run(64, N, nacc, 0);
  for (int N : {48, 96, 192}) run(128, N, 1, 1);
  for (int mode : {2, 3, 4})
    for (int N : {48, 96, 192}) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) { kern<<<1, 128, smem>>>(128, N, 1, total, mode, dC); cudaDeviceSynchronize(); cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); }
      printf("nested loop, %s, N=%3d: %.1f cycles per MMA\n", mode == 2 ? "elect.sync per instruction" : mode == 3 ? "leader predicate passed in" : "one elected thread runs the loop", N, (double)c / total);
    }
  printf("OK\n");
  return 0;
}
