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
  CUtensorMap o[2];       // output rows [nrows, Cout] hi/lo, box {tile_n channels, 128 rows} (TcArgs.tma_out)
};

struct TcArgs {
  int C[2];               // channels of the (up to two, concatenated) main-phase sources; C[1] may be 0
  int RC[2];              // residual-phase sources (both 0 => no residual conv)
  int T;                  // tap blocks computed by the main phase (packed-weight taps tap0 .. tap0+T-1)
  int tap0;
  int n_out;              // outputs per GEMM row: 1 (conv, downsample) or 2 (transposed conv: rows 2m and 2m+1)
  int nt[2];              // taps contributing to output o
  int tap_blk[2][5];      // [o][i] tap block
  int tap_shift[2][5];    // [o][i] row shift: out_o[l] += Y_blk[l + shift]
  int out_ldiv;           // 2 for the stride-2 conv: only rows l % 2 == 0 emit output row l/2
  int out_L, out_lmul;    // output row = b*out_L + (l / out_ldiv)*out_lmul + o
  int Cout;
  int nrows;              // B * Lrows (GEMM rows)
  int Lrows, log2L;       // GEMM rows per sample (= input positions)
  int samples_per_tile;   // 128 / Lrows
  // epilogue
  const float* bias;
  const float* gn_gamma; const float* gn_beta; int cg;
  const float* temb; int temb_stride;
  const float* temb2;     // per-step time term shared by the batch (may be null)
  const float* resB;
  const float* res_f32;                              // identity residual, fp32 [nrows, Cout]
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;   // identity residual, bf16 hi/lo
  const float* headW; const float* headB; int head_dim; float* head_out;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  // auxiliary output: tap block `aux_blk` (>= 0) is NOT part of the conv sum; it is a 1x1 conv of the same input rows (the block's
  // residual projection packed as one more "tap"), written as fp32 rows  aux_out[row, Cout] = Y_aux + aux_bias  for the chain kernel
  float* aux_out; const float* aux_bias; int aux_blk;
  int tile_n;             // column-tile width: 64, 32 or 16 (tc_pick_tile_n)
  int cluster_n;          // set by launch_conv_tc: CTAs along N sharing one GroupNorm group
  int cluster_l;          // set by launch_conv_tc: cluster size along N (activation-tile multicast), multiple of cluster_n
  int cluster_m;          // set by tc_configure: cluster size along M (row tiles sharing a column tile): every CTA loads 1/cluster_m of the WEIGHT stage and
                          // multicasts it to the others (large batch: the weight tile re-read by every row tile is most of the L2 traffic); 1 = off
  int ring;               // set by launch_conv_tc: bytes of the operand ring in dynamic shared memory
  int stages;             // set by launch_conv_tc: ring stages = min(ring / stage bytes, 8)
  int concat;             // set by launch_conv_tc (bf16x3): hi x [W_hi | W_lo] as ONE MMA of N = 2*T*tile_n; the hi*lo products get their own TMEM block
  int tma_out;            // set by the caller with maps.o: the tile leaves through a swizzled staging buffer and one TMA store per plane instead of per-thread
                          // 8-byte global stores (a warp's store instruction touched 32 different lines: 5 % of an iteration at B = 256, 19 % at B = 4096).
                          // Only for n_out == 1, out_ldiv == 1, no fused head (output row == GEMM row)
  int w_hint;             // set by launch_conv_tc (B2P_TC_WHINT): 1 (default) = weight TMA loads carry an L2 evict_last hint (-0.5 % per iteration), 2 = and activation loads evict_first
  int dbg;                // developer bisect switch (B2P_TC_DBG): 1 = skip the TMA/MMA main loop, 2 = skip the epilogue math
};

int tc_make_act_map(CUtensorMap* m, const void* base, int B, int L, int C, int box_l, int lstride, int box_b);
int tc_make_weight_map(CUtensorMap* m, const void* base, int taps, int Cout, int K, int box_taps, int box_n);
int tc_make_out_map(CUtensorMap* m, const void* base, int nrows, int Cout, int tile_n);
int tc_pick_tile_n(int nrows, int Cout, bool has_head);
int tc_configure(TcArgs& a);
int launch_conv_tc(const TcMaps& maps, const TcArgs& a, int nsplit, cudaStream_t s);
const char* tc_last_error();   // detail of the last B2P_ERR_INVALID_ARG returned by launch_conv_tc on this thread

}  // namespace b2p
