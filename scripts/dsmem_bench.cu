// Developer micro-benchmark: cost of exchanging a [128 x 64] fp32 partial tile between the CTAs of a cluster
// (reduce-scatter by rows) via DSMEM push, DSMEM pull, or global memory (L2).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// mode 0: nothing but 2 cluster syncs; 1: push (st.shared::cluster); 2: pull (ld.shared::cluster); 3: global write + read
__global__ void __launch_bounds__(512, 1) k_xchg(float* gws, float* out, int mode, int reps) {
  extern __shared__ __align__(16) float sm[];   // [KS][128/KS][64] receive buffer (push) or own tile [128][64] (pull)
  const int KS = gridDim.z, rank = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (warp & 3) * 32 + lane, col0 = (warp >> 2) * 16;
  const int rpo = 128 / KS, rl = r % rpo, owner = r / rpo;
  float v[16];
  for (int c = 0; c < 16; ++c) v[c] = (float)(threadIdx.x + c + rank);
  float acc[16] = {0};
  for (int it = 0; it < reps; ++it) {
    csync();
    if (mode == 1) {
      uint32_t la = su32(sm + ((size_t)rank * rpo + rl) * 64 + col0), ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(owner));
      for (int c = 0; c < 16; c += 4)
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra + c * 4), "f"(v[c]), "f"(v[c + 1]), "f"(v[c + 2]), "f"(v[c + 3]) : "memory");
    } else if (mode == 2) {
      float4* p = reinterpret_cast<float4*>(sm + (size_t)r * 64 + col0);
      for (int c = 0; c < 4; ++c) p[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    } else if (mode == 3) {
      size_t tile = (size_t)blockIdx.x * KS + rank;
      float4* p = reinterpret_cast<float4*>(gws + (tile * 128 + r) * 64 + col0);
      for (int c = 0; c < 4; ++c) p[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    csync();
    if (mode == 1) {
      for (int s = 0; s < KS; ++s) {
        const float4* q = reinterpret_cast<const float4*>(sm + ((size_t)s * rpo + rl) * 64 + col0);
        for (int c = 0; c < 4; ++c) { float4 t = q[c]; acc[4 * c] += t.x; acc[4 * c + 1] += t.y; acc[4 * c + 2] += t.z; acc[4 * c + 3] += t.w; }
      }
    } else if (mode == 2) {
      if (owner == rank) {
        for (int s = 0; s < KS; ++s) {
          uint32_t la = su32(sm + (size_t)r * 64 + col0), ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(s));
          for (int c = 0; c < 4; ++c) {
            float4 t;
            asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(ra + c * 16));
            acc[4 * c] += t.x; acc[4 * c + 1] += t.y; acc[4 * c + 2] += t.z; acc[4 * c + 3] += t.w;
          }
        }
      }
    } else if (mode == 3) {
      if (owner == rank) {
        for (int s = 0; s < KS; ++s) {
          size_t tile = (size_t)blockIdx.x * KS + s;
          const float4* q = reinterpret_cast<const float4*>(gws + (tile * 128 + r) * 64 + col0);
          for (int c = 0; c < 4; ++c) { float4 t = __ldcg(q + c); acc[4 * c] += t.x; acc[4 * c + 1] += t.y; acc[4 * c + 2] += t.z; acc[4 * c + 3] += t.w; }
        }
      }
    }
  }
  csync();
  float s = 0; for (int c = 0; c < 16; ++c) s += acc[c];
  if (s == 123.456f) out[0] = s;
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  float *gws, *out; cudaMalloc(&gws, 256 * 128 * 64 * 4); cudaMalloc(&out, 4);
  cudaFuncSetAttribute(k_xchg, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[4] = {"2 cluster syncs only", "DSMEM push + local sum", "DSMEM pull (owner rows)", "global write + L2 read"};
  for (int ks : {2, 4, 8}) for (int mode = 0; mode < 4; ++mode) {
    int reps = 200;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128 / ks, 1, ks); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 32 * 1024; cfg.stream = st;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = ks; cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchKernelEx(&cfg, k_xchg, gws, out, mode, reps); cudaStreamSynchronize(st);
    cudaEventRecord(a, st); cudaLaunchKernelEx(&cfg, k_xchg, gws, out, mode, reps); cudaEventRecord(b, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("KS=%d %-26s: %.2f us per exchange (%s)\n", ks, names[mode], ms * 1000.f / reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
