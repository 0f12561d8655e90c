// Developer test: tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st from registers) against the same product
// with A in shared memory and against the host.  Establishes the TMEM layout of a bf16 A operand (row = lane, two bf16 per 32-bit
// column, K ascending along columns) and times both forms: N = 96 / 48 per instruction, 32 K steps (a 512-channel layer's main loop).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
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
__device__ __forceinline__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
// warp-uniform issue: the WHOLE warp runs the loop (operands stay in uniform registers), one elected lane issues the instruction
__device__ __forceinline__ void mma_ss_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t swz(int r, int c8) { return (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4)); }

// A [128][K] bf16 row-major, B [N][K] bf16 row-major (K-major), D [128][N] fp32.  K = 64 * chunks.  mode 0: SS, 1: TS.
template <int N>
__global__ void __launch_bounds__(128) kern(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int chunks, int mode, int reps, long long* cycles, int nacc = 1, int elect = 0) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // chunks x 16 KB
  uint8_t* sB = smem + chunks * 16384;      // chunks x N*128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, K = chunks * 64;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stage A and B in swizzled K-major smem
  for (int i = tid; i < chunks * 128 * 8; i += 128) {
    const int ch = i / 1024, r = (i >> 3) & 127, c8 = i & 7;
    *(uint4*)(sA + ch * 16384 + swz(r, c8)) = *(const uint4*)(A + (size_t)r * K + ch * 64 + c8 * 8);
  }
  for (int i = tid; i < chunks * N * 8; i += 128) {
    const int ch = i / (N * 8), r = (i >> 3) % N, c8 = i & 7;
    *(uint4*)(sB + ch * N * 128 + swz(r, c8)) = *(const uint4*)(B + (size_t)r * K + ch * 64 + c8 * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  const uint32_t a_col0 = 256;              // A operand columns: chunk c at [256 + 32 c, +32)  (<= 8 chunks)
  if (mode == 1) {
    // row r = lane of quadrant `warp`: 64 channels of a chunk = 128 B = 32 columns of packed bf16 pairs
    const int r = warp * 32 + lane;
    for (int ch = 0; ch < chunks; ++ch) {
      uint32_t v[32];
      const uint4* src = (const uint4*)(A + (size_t)r * K + ch * 64);
#pragma unroll
      for (int q = 0; q < 8; ++q) { uint4 t = src[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
      const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + a_col0 + ch * 32;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                     "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                     "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t0 = 0, t1 = 0;
  if (elect && warp == 0) {
    const uint32_t id = idesc_n(N);
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int ch = 0; ch < chunks; ++ch) {
        const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)), bd = umma_desc(smem_u32(sB + ch * N * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (rep == 0 && ch == 0 && k == 0) ? 0u : 1u;
          if (mode == 0) mma_ss_elect(tb, ad + 2 * k, bd + 2 * k, id, acc);
          else mma_ts_elect(tb, tb + a_col0 + ch * 32 + k * 8, bd + 2 * k, id, acc);
        }
      }
    }
    if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
  }
  if (!elect && tid == 0) {
    const uint32_t id = idesc_n(N);
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int ch = 0; ch < chunks; ++ch) {
        const uint64_t ad = umma_desc(smem_u32(sA + ch * 16384)), bd = umma_desc(smem_u32(sB + ch * N * 128));
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (rep == 0 && ch == 0 && k == 0) ? 0u : 1u;
          const uint32_t dcol = tb + ((ch * 4 + k) & (nacc - 1)) * N;    // independent accumulators (timing experiment; nacc * N <= 256)
          if (mode == 0) mma_ss(dcol, ad + 2 * k, bd + 2 * k, id, acc);
          else mma_ts(dcol, tb + a_col0 + ch * 32 + k * 8, bd + 2 * k, id, acc);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(&bar)) : "memory");
  if (tid == 0) { t1 = clock64(); *cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)r * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

template <int N> int run(int chunks) {
  const int K = chunks * 64;
  std::vector<uint16_t> hA(128 * K), hB(N * K);
  srand(1);
  for (auto& v : hA) v = f2bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : hB) v = f2bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<float> ref(128 * N, 0.f);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)bf2f(hA[r * K + k]) * bf2f(hB[n * K + k]); ref[r * N + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = chunks * 16384 + chunks * N * 128 + 2048;
  cudaFuncSetAttribute(kern<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad = 0;
  for (int mode = 0; mode < 2; ++mode) {
    std::vector<float> out(128 * N);
    kern<N><<<1, 128, smem>>>(dA, dB, dD, chunks, mode, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs(out[i] - ref[i]));
    long long c1, c8;
    cudaMemcpy(&c1, dC, 8, cudaMemcpyDeviceToHost);
    kern<N><<<1, 128, smem>>>(dA, dB, dD, chunks, mode, 8, dC);
    cudaDeviceSynchronize();
    cudaMemcpy(&c8, dC, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d K=%4d %s: %s max err %.3e; cycles per MMA (K=16): %.1f  (8 passes: %lld cycles for %d MMAs)\n", N, K, mode ? "A in TMEM (TS)" : "A in smem (SS)",
           cudaGetErrorString(e), err, (double)c8 / (8.0 * chunks * 4), c8, 8 * chunks * 4);
    if (e != cudaSuccess || err > 1e-2) bad = 1;
    {
      kern<N><<<1, 128, smem>>>(dA, dB, dD, chunks, mode, 1, dC, 1, 1);
      cudaError_t e2 = cudaDeviceSynchronize();
      cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
      double err2 = 0; for (size_t i = 0; i < out.size(); ++i) err2 = fmax(err2, fabs(out[i] - ref[i]));
      kern<N><<<1, 128, smem>>>(dA, dB, dD, chunks, mode, 8, dC, 1, 1);
      cudaDeviceSynchronize();
      cudaMemcpy(&c8, dC, 8, cudaMemcpyDeviceToHost);
      printf("        warp-uniform issue (elect.sync): %s err %.3e, cycles per MMA %.1f\n", cudaGetErrorString(e2), err2, (double)c8 / (8.0 * chunks * 4));
      if (e2 != cudaSuccess || err2 > 1e-2) bad = 1;
    }
    for (int nacc = 2; nacc <= 4 && nacc * N <= 256; nacc *= 2) {
      kern<N><<<1, 128, smem>>>(dA, dB, dD, chunks, mode, 8, dC, nacc);
      cudaDeviceSynchronize();
      cudaMemcpy(&c8, dC, 8, cudaMemcpyDeviceToHost);
      printf("        %d independent accumulators: cycles per MMA %.1f\n", nacc, (double)c8 / (8.0 * chunks * 4));
    }
  }
  return bad;
}

int main() {
  int bad = 0;
  bad |= run<16>(8); bad |= run<48>(8); bad |= run<96>(8); bad |= run<192>(4); bad |= run<256>(4);
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
