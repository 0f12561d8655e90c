// Developer probe for the encoder's 3x3 convolution on tcgen05: can ONE shared-memory halo patch serve all nine taps?
// The patch is [PH = 18][PW = 10] pixels x 64 channels (128 B per pixel), stored the way TMA stores a {64, PW, PH} box with SWIZZLE_128B:
// 16-byte chunk j of the pixel at linear index p sits at p * 128 + ((j ^ (p & 7)) << 4).  The A operand of tap (dy, dx) for a tile of 16 rows x 8
// columns of output pixels is then rows m = r * 8 + c -> patch pixel (r + dy, c + dx): 8-row groups PW * 128 B apart (SBO), starting
// (dy * PW + dx) * 128 B into the patch — a start address that is NOT aligned to the 1024-byte swizzle pattern.  This program checks the
// product against the host for every tap with the descriptor's base-offset field (bits 49..51) set to 0 and to (start >> 7) & 7.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int PW = 10, PH = 18;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_n(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// patch: [PH*PW][64] bf16 (global, plain), W: [9][64 out][64 in] bf16, D: [128][64] fp32
__global__ void __launch_bounds__(128) kern(const uint16_t* patch, const uint16_t* W, float* D, int use_base_off) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // PH*PW*128 = 23040 B -> 23 KB (padded to 24 KB)
  uint8_t* sB = smem + 24 * 1024;     // 9 taps x 8 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < PH * PW * 8; i += 128) {
    const int p = i >> 3, j = i & 7;
    *(uint4*)(sA + p * 128 + ((j ^ (p & 7)) << 4)) = *(const uint4*)(patch + (size_t)p * 64 + j * 8);
  }
  for (int i = tid; i < 9 * 64 * 8; i += 128) {
    const int t = i / 512, n = (i >> 3) & 63, j = i & 7;
    *(uint4*)(sB + t * 8192 + n * 128 + ((j ^ (n & 7)) << 4)) = *(const uint4*)(W + ((size_t)t * 64 + n) * 64 + j * 8);
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  if (tid == 0) {
    const uint32_t id = idesc_n(64);
    int first = 1;
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t % 3;
      const uint32_t a0 = smem_u32(sA) + (uint32_t)(dy * PW + dx) * 128u;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_desc(a0 + k * 32, PW * 128, use_base_off ? (a0 >> 7) & 7 : 0);
        const uint64_t bd = umma_desc(smem_u32(sB) + t * 8192 + k * 32, 1024, 0);
        const uint32_t acc = first ? 0u : 1u;
        first = 0;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tb), "l"(ad), "l"(bd), "r"(id), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)r * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory");
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  std::vector<uint16_t> hp(PH * PW * 64), hw(9 * 64 * 64);
  srand(3);
  for (auto& v : hp) v = f2bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : hw) v = f2bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<float> ref(128 * 64, 0.f);
  for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) for (int n = 0; n < 64; ++n) {
    double s = 0;
    for (int t = 0; t < 9; ++t) { const int p = (r + t / 3) * PW + c + t % 3; for (int k = 0; k < 64; ++k) s += (double)bf2f(hp[p * 64 + k]) * bf2f(hw[(t * 64 + n) * 64 + k]); }
    ref[(r * 8 + c) * 64 + n] = (float)s;
  }
  uint16_t *dp, *dw; float* dd;
  cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
  cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 24 * 1024 + 9 * 8192 + 2048;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int bo = 0; bo < 2; ++bo) {
    cudaMemset(dd, 0, 128 * 64 * 4);
    kern<<<1, 128, smem>>>(dp, dw, dd, bo);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 64);
    cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; int bad_rows = 0;
    for (int m = 0; m < 128; ++m) { double re = 0; for (int n = 0; n < 64; ++n) re = fmax(re, fabs(out[m * 64 + n] - ref[m * 64 + n])); if (re > 1e-2) ++bad_rows; err = fmax(err, re); }
    printf("base_offset field %s: %s, max err %.3e, rows off %d / 128\n", bo ? "= (start >> 7) & 7" : "= 0", cudaGetErrorString(e), err, bad_rows);
  }
  return 0;
}
