// Host/device-shared declarations of the tcgen05 conv path (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace b2p {

struct TcMaps {
  CUtensorMap a[2][2];    // main-phase activation sources [source][hi/lo]
  CUtensorMap r[2][2];    // residual-phase activation sources
  CUtensorMap w[2];       // packed conv weights [taps*Cout, Cin]   hi/lo
  CUtensorMap rw[2];      // residual 1x1 weights [Cout, RCin]      hi/lo
};

struct TcArgs {
  int C[2];               // channels of the (up to two, concatenated) main-phase sources; C[1] may be 0
  int RC[2];              // residual-phase sources (both 0 => no residual conv)
  int ntaps;
  int tap_l0[2][5];       // [parity][tap] first input position of the row-shifted box
  int tap_w[2][5];        // [parity][tap] tap index into the packed weights
  int Cout;
  int nrows;              // B * Lrows
  int Lrows, log2L;       // rows per sample in this GEMM
  int samples_per_tile;   // 128 / Lrows
  // epilogue
  const float* bias;
  const float* gn_gamma; const float* gn_beta; int cg;
  const float* temb; int temb_stride;
  const float* resB;
  const float* res_f32;                              // identity residual, fp32 [nrows, Cout]
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;   // identity residual, bf16 hi/lo
  const float* headW; const float* headB; int head_dim; float* head_out;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int out_L, out_lstride, out_loff0, out_loff1;      // output row = b*out_L + l*out_lstride + out_loff[parity]
};

int tc_make_act_map(CUtensorMap* m, const void* base, int B, int L, int C, int box_l, int lstride, int box_b);
int tc_make_weight_map(CUtensorMap* m, const void* base, int rows, int K);
int launch_conv_tc(const TcMaps& maps, const TcArgs& a, int nsplit, int nparity, cudaStream_t s);

}  // namespace b2p
