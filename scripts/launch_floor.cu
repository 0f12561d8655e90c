// Developer micro-benchmark: per-launch cost of dependent kernels inside a CUDA graph on B200, for the launch shapes the
// tcgen05 conv kernel uses (large dynamic smem, 512 threads, ~2 KB of parameters, TMEM allocation).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
struct Big { char pad[1800]; };
__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_params(const __grid_constant__ Big b, int* p) { if (p && threadIdx.x == 9999) *p = b.pad[3]; }
__global__ void k_tmem(int* p, uint32_t cols) {
  __shared__ uint32_t base;
  extern __shared__ char dyn[];
  if (threadIdx.x / 32 == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x / 32 == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
  if (p && threadIdx.x == 9999) *p = dyn[0];
}
__global__ void k_cluster(int* p, int nsync) {
  for (int i = 0; i < nsync; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (p && threadIdx.x == 9999) *p = 1;
}
static void launch_cluster(int gx, int gz, int threads, int smem, int nsync, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, 1, gz); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = gz;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int* p = nullptr;
  cudaLaunchKernelEx(&cfg, k_cluster, p, nsync);
}
template <typename F> float time_graph(F launch, int n, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t e;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) launch();
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&e, g, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
  cudaEventRecord(a, s); cudaGraphLaunch(e, s); cudaEventRecord(b, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
  return ms * 1000.f / n;
}
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int n = 2000; Big big{};
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_params, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  for (int gz : {1, 2, 4, 8}) {
    printf("cluster z=%d\n", gz);
    printf("  16 clusters, 512thr 226KB, 0 syncs : %.2f us/launch\n", time_graph([&] { launch_cluster(16, gz, 512, 226 * 1024, 0, s); }, n, s));
    printf("  16 clusters, 512thr 226KB, 2 syncs : %.2f us/launch\n", time_graph([&] { launch_cluster(16, gz, 512, 226 * 1024, 2, s); }, n, s));
    printf("  16 clusters, 512thr 0KB,   2 syncs : %.2f us/launch\n", time_graph([&] { launch_cluster(16, gz, 512, 0, 2, s); }, n, s));
    printf("  1 cluster,   512thr 226KB, 2 syncs : %.2f us/launch\n", time_graph([&] { launch_cluster(1, gz, 512, 226 * 1024, 2, s); }, n, s));
  }
  for (int grid : {1, 32, 148}) {
    printf("grid %d\n", grid);
    printf("  empty 128thr 0smem         : %.2f us/launch\n", time_graph([&] { k_empty<<<grid, 128, 0, s>>>(nullptr); }, n, s));
    printf("  empty 512thr 0smem         : %.2f us/launch\n", time_graph([&] { k_empty<<<grid, 512, 0, s>>>(nullptr); }, n, s));
    printf("  empty 512thr 226KB smem    : %.2f us/launch\n", time_graph([&] { k_empty<<<grid, 512, 226 * 1024, s>>>(nullptr); }, n, s));
    printf("  +1.8KB params              : %.2f us/launch\n", time_graph([&] { k_params<<<grid, 512, 226 * 1024, s>>>(big, nullptr); }, n, s));
    printf("  tmem alloc 128 cols 226KB  : %.2f us/launch\n", time_graph([&] { k_tmem<<<grid, 512, 226 * 1024, s>>>(nullptr, 128); }, n, s));
    printf("  tmem alloc 512 cols 226KB  : %.2f us/launch\n", time_graph([&] { k_tmem<<<grid, 512, 226 * 1024, s>>>(nullptr, 512); }, n, s));
    printf("  tmem alloc 512 cols 0smem  : %.2f us/launch\n", time_graph([&] { k_tmem<<<grid, 512, 0, s>>>(nullptr, 512); }, n, s));
    printf("  alternate 0smem/226KB      : %.2f us/launch\n", time_graph([&] { static int i = 0; if (i++ & 1) k_empty<<<grid, 512, 226 * 1024, s>>>(nullptr); else k_empty<<<grid, 256, 0, s>>>(nullptr); }, n, s));
  }
  return 0;
}
