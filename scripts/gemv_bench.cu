// Developer harness: chains of small-batch GEMV conv layers with the denoiser's layer shapes at batch 1 inside a CUDA graph,
// with per-stage clocks from the kernel (B2P_GV_TRACE).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o scripts/gemv_bench scripts/gemv_bench.cu
#define B2P_GV_TRACE 1
#include "../autonomous_driving_with_diffusion_model_b200/csrc/conv_gemv.cu"
#include <cstdio>
#include <vector>
using namespace b2p;

struct Shape { const char* name; int C0, C1, Cout, L, taps, jmin, jmax, gn, res; };

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1;
  cudaStream_t s; cudaStreamCreate(&s);
  const size_t wmax = 5u * 1024 * 512 + 1024, amax = 64 * 16 * 1024;
  float *W, *RW, *act[3], *vec;
  cudaMalloc(&W, wmax * 4 * 2); cudaMalloc(&RW, 1024 * 512 * 4);
  for (auto& a : act) { cudaMalloc(&a, amax * 4); cudaMemset(a, 0, amax * 4); }
  cudaMalloc(&vec, 4096 * 4);
  { std::vector<float> h(wmax * 2); for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) >> 20 & 255) / 25600.f - 0.005f; cudaMemcpy(W, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(RW, h.data(), 1024 * 512 * 4, cudaMemcpyHostToDevice); cudaMemcpy(vec, h.data(), 4096 * 4, cudaMemcpyHostToDevice); }
  const Shape shapes[] = {
      {"64->64 L16 k5 gn", 64, 0, 64, 16, 5, 0, 4, 1, 0},      {"128->128 L8 k5 gn", 128, 0, 128, 8, 5, 0, 4, 1, 0},
      {"256->256 L4 k5 gn", 256, 0, 256, 4, 5, 0, 4, 1, 0},    {"512->512 L2 k5(3) gn", 512, 0, 512, 2, 5, 1, 3, 1, 0},
      {"512->512 L2 k5(3) gn +id res", 512, 0, 512, 2, 5, 1, 3, 1, 2}, {"(256+256)->256 L2 gn", 256, 256, 256, 2, 5, 1, 3, 1, 0},
      {"256->256 L2 gn + res 1x1 512", 256, 0, 256, 2, 5, 1, 3, 1, 1}, {"64->64 L16 k3 plain (no gn)", 64, 0, 64, 16, 3, 0, 2, 0, 0},
  };
  const int n = 400;
  for (const Shape& sh : shapes) {
    cudaGraph_t g; cudaGraphExec_t e;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < n; ++i) {
      ConvArgs a{};
      a.x0 = act[i & 1]; a.C0 = sh.C0; a.x1 = sh.C1 ? act[2] : nullptr; a.C1 = sh.C1;
      a.Lin = a.Lout = sh.L; a.log2Lout = 0; a.nrows = B * sh.L; a.Cout = sh.Cout; a.taps = sh.taps; a.jmin = sh.jmin; a.jmax = sh.jmax;
      a.stride = 1; a.pad = sh.taps / 2; a.Wk = W + (i % 2) * wmax; a.bias = vec;
      if (sh.gn) { a.gn_gamma = vec + 1024; a.gn_beta = vec + 2048; a.cg = sh.Cout / 8; }
      a.temb2 = vec + 3072;
      if (sh.res == 2) a.res_id = act[2];
      if (sh.res == 1) { a.rx0 = act[2]; a.RC0 = 512; a.resW = RW; a.resWk = RW; a.resB = vec + 512; }
      a.out = act[(i + 1) & 1];
      int rc = launch_conv_gemv(a, s);
      if (rc) { printf("launch failed %d\n", rc); return 1; }
    }
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&e, g, 0);
    cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
    cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
    cudaEventRecord(ea, s); cudaGraphLaunch(e, s); cudaEventRecord(eb, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, ea, eb);
    cudaError_t err = cudaGetLastError();
    static unsigned long long tr[8192 * 8];
    cudaMemcpyFromSymbol(tr, gv_trace, sizeof(tr));
    static int base = 0;   // launch ids keep counting across shapes (gv_trace_launch)
    double st[6] = {0, 0, 0, 0, 0, 0}, span = 0;
    int cnt = 0;
    for (int i = 100; i < n - 1; ++i) {
      const unsigned long long* t = tr + (size_t)((base + i) & 8191) * 8;
      const unsigned long long* tn = tr + (size_t)((base + i + 1) & 8191) * 8;
      for (int k = 0; k < 5; ++k) st[k] += (double)(t[k + 1] - t[k]);
      span += (double)(tn[7] - t[7]);   // end-to-end distance between consecutive kernels (globaltimer, ns)
      ++cnt;
    }
    base += n;
    printf("%-34s B=%d: %.2f us/kernel (%s) | cycles: prologue %.0f wait %.0f X+weights %.0f dot %.0f combine+GN %.0f | end-to-end %.0f ns\n", sh.name, B,
           ms * 1000.f / n, cudaGetErrorString(err), st[0] / cnt, st[1] / cnt, st[2] / cnt, st[3] / cnt, st[4] / cnt, span / cnt);
    cudaGraphExecDestroy(e); cudaGraphDestroy(g);
  }
  return 0;
}
