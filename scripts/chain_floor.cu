// Developer micro-benchmark: cost per kernel of a chain of DATA-DEPENDENT kernels inside a CUDA graph on B200 — the floor
// under one denoiser layer.  Kernel i reads what kernel i-1 wrote (one global load per thread), optionally chases
// `hops` further dependent L2 loads, optionally syncs a cluster, and writes its output.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/chain_floor scripts/chain_floor.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void k_chain(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ chase, int hops, int nsync,
                        int trigger_early, int csync) {
  extern __shared__ float dyn[];
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = in[i];
  int idx = (int)v & 1023;
  for (int h = 0; h < hops; ++h) idx = chase[idx];     // dependent L2 round trips
  v += (float)idx;
  for (int s = 0; s < nsync; ++s) {
    dyn[threadIdx.x] = v;
    __syncthreads();
    v = dyn[(threadIdx.x + 1) % blockDim.x] * 0.5f + 1.f;
    __syncthreads();
  }
  for (int s = 0; s < csync; ++s) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  out[i] = v * 0.25f;
}

static void launch(const float* in, float* out, const int* chase, int grid, int threads, int smem, int hops, int nsync, int pdl,
                   int early, int cluster, int csync, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) { attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  if (cluster > 1) { attr[n].id = cudaLaunchAttributeClusterDimension; attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1; ++n; }
  cfg.attrs = attr; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, k_chain, in, out, chase, hops, nsync, early, csync);
}

template <typename F> float time_graph(F body, int n, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t e;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) body(i);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&e, g, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
  cudaEventRecord(a, s); cudaGraphLaunch(e, s); cudaEventRecord(b, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
  cudaGraphExecDestroy(e); cudaGraphDestroy(g);
  return ms * 1000.f / n;
}

int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  const int n = 2000, maxel = 148 * 4 * 1024;
  float *a, *b; int* chase;
  cudaMalloc(&a, maxel * 4); cudaMalloc(&b, maxel * 4); cudaMalloc(&chase, 1024 * 4);
  cudaMemset(a, 0, maxel * 4); cudaMemset(b, 0, maxel * 4);
  int h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(chase, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int grid, threads, smem; };
  for (Cfg c : {Cfg{8, 256, 4096}, Cfg{64, 256, 4096}, Cfg{64, 1024, 4096}, Cfg{64, 256, 100 * 1024}, Cfg{128, 512, 100 * 1024}, Cfg{296, 512, 100 * 1024}}) {
    printf("grid %d x %d thr, %d KB smem\n", c.grid, c.threads, c.smem / 1024);
    auto run = [&](const char* name, int hops, int nsync, int pdl, int early, int cluster, int csync) {
      float us = time_graph([&](int i) { launch(i & 1 ? b : a, i & 1 ? a : b, chase, c.grid, c.threads, c.smem, hops, nsync, pdl, early, cluster, csync, s); }, n, s);
      printf("  %-46s: %.2f us/kernel\n", name, us);
    };
    run("plain launches", 0, 0, 0, 0, 1, 0);
    run("PDL, trigger after wait", 0, 0, 1, 0, 1, 0);
    run("PDL, trigger at start", 0, 0, 1, 1, 1, 0);
    run("PDL early + 1 L2 hop", 1, 0, 1, 1, 1, 0);
    run("PDL early + 2 L2 hops", 2, 0, 1, 1, 1, 0);
    run("PDL early + 4 L2 hops", 4, 0, 1, 1, 1, 0);
    run("PDL early + 3 smem syncs", 0, 3, 1, 1, 1, 0);
    run("PDL early, cluster 8, no sync", 0, 0, 1, 1, 8, 0);
    run("PDL early, cluster 8, 1 cluster sync", 0, 0, 1, 1, 8, 1);
    run("PDL early, cluster 8, 2 cluster syncs", 0, 0, 1, 1, 8, 2);
    run("PDL early, cluster 2, 1 cluster sync", 0, 0, 1, 1, 2, 1);
    run("plain, cluster 8, 1 cluster sync", 0, 0, 0, 0, 8, 1);
  }
  return 0;
}
