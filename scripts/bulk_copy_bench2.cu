// Developer micro-benchmark, second part: the 80 KB weight copy of chain64.cu takes 5-8 k cycles inside the kernel and 1.5 k in isolation
// (bulk_copy_bench.cu).  This program adds the chain kernel's launch conditions one at a time: 544 threads and 204 KB of dynamic shared memory with the
// destination 96 KB in, a 512-column TMEM allocation, the programmatic-stream-serialization launch attribute with a preceding kernel, an L2 access-policy
// window, a cold source (other traffic between uses), and a copy issued while the other warps sleep on an mbarrier / sit in bar.sync.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void dummy(float* p) { if (threadIdx.x == 0 && p) p[blockIdx.x] = 1.f; }
__global__ void touch(const uint4* p, size_t n, uint4* sink) {   // stream other data through L2
  uint4 a = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = p[i]; a.x ^= v.x; }
  if (a.x == 0x12345u) *sink = a;
}
// mode bits: 1 = allocate 512 TMEM columns, 2 = griddepcontrol (launch_dependents + wait), 4 = the other warps wait in bar.sync while the copy flies,
//            8 = the other warps sleep on an mbarrier with a suspend-time hint
__global__ void __launch_bounds__(544, 1) kern(const uint8_t* src, int bytes, int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* dst = smem + 96 * 1024;
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if ((mode & 1) && warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (mode & 2) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  long long tot = 0;
  uint32_t par = 0;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    long long t0 = clock64();
    if (warp == 16) {
      if ((tid & 31) == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        for (int i = 0; i < 10; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + i * 8192)), "l"(src + i * 8192),
                       "r"(bytes / 10), "r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(&bar)), "r"(par) : "memory");
        tot += clock64() - t0;
        if (mode & 8) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
    } else if (mode & 8) {
      asm volatile("{\n.reg .pred p;\nS_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra E_%=;\nbra S_%=;\nE_%=:\n}" ::"r"(smem_u32(&bar2)), "r"(par), "r"(1000000u) : "memory");
    }
    par ^= 1u;
    if (mode & 4) __syncthreads();
  }
  __syncthreads();
  if (tid == 16 * 32) out[blockIdx.x] = tot / reps;
  if ((mode & 1) && warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
  const int bytes = 80 * 1024;
  uint8_t *src, *big; long long* out; float* dp; uint4* sink;
  cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  const size_t bigb = 96u << 20;
  cudaMalloc(&big, bigb); cudaMemset(big, 2, bigb);
  cudaMalloc(&out, 148 * 8); cudaMalloc(&dp, 4096); cudaMalloc(&sink, 16);
  const int smem = 204 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaStream_t s; cudaStreamCreate(&s);
  auto run = [&](const char* name, int mode, int reps, bool pdl, bool window, bool cold) {
    long long h[32]; long long sum = 0;
    for (int it = 0; it < 3; ++it) {
      if (cold) touch<<<592, 256, 0, s>>>((const uint4*)big, bigb / 16, sink);
      dummy<<<148, 128, 0, s>>>(dp);
      cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(32); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute attr[2]; int na = 0;
      if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
      if (window) { attr[na].id = cudaLaunchAttributeAccessPolicyWindow; attr[na].val.accessPolicyWindow.base_ptr = big; attr[na].val.accessPolicyWindow.num_bytes = 62u << 20;
                    attr[na].val.accessPolicyWindow.hitRatio = 1.f; attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming; ++na; }
      cfg.attrs = attr; cfg.numAttrs = na;
      cudaLaunchKernelEx(&cfg, kern, (const uint8_t*)src, bytes, mode, reps, out);
      cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, out, 32 * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 32; ++i) sum += h[i];
    printf("%-100s %6lld cycles\n", name, sum / 32);
  };
  run("544 threads, 204 KB smem, destination 96 KB in, 8 copies in a row (hot)", 0, 8, false, false, false);
  run("  + 512 TMEM columns allocated", 1, 8, false, false, false);
  run("  + TMEM + griddepcontrol, launched with programmatic stream serialization", 3, 8, true, false, false);
  run("  + TMEM + PDL + L2 access-policy window (62 MB persisting elsewhere)", 3, 8, true, true, false);
  run("  + TMEM + PDL + other warps in bar.sync while the copy flies", 7, 8, true, false, false);
  run("  + TMEM + PDL + other warps asleep on an mbarrier (suspend-time hint)", 11, 8, true, false, false);
  run("  ONE copy per launch, source cold (96 MB streamed through L2 before the launch)", 3, 1, true, false, true);
  run("  ONE copy per launch, source cold, + window", 3, 1, true, true, true);
  run("  ONE copy per launch, source cold, others asleep on an mbarrier", 11, 1, true, false, true);
  return 0;
}
