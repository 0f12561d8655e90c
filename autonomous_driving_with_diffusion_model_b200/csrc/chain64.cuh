// Host/device-shared declarations of the row-owned 64-channel chain kernel (chain64.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "sched_math.cuh"

namespace b2p {

constexpr int CH_NS = 8;          // most trajectories one CTA can own (8 x 16 positions = one 128-row MMA tile)
constexpr int CH_MAXOPS = 12;

enum { CH_CONV = 0, CH_DOWN = 1, CH_UP = 2 };                               // ChainOp.kind
enum { CH_RES_NONE = 0, CH_RES_SMEM = 1, CH_RES_XPROJ = 2, CH_RES_F32 = 3 };  // ChainOp.res_kind
enum { CH_OUT_GLOBAL = -1, CH_OUT_HEAD = -2 };                              // ChainOp.out_buf
enum { CH_IN_IM2COL = -1 };                                                 // ChainOp.in_buf

struct ChainOp {
  int kind;               // CH_CONV: k taps, stride 1;  CH_DOWN: Conv1d(k3, s2, p1);  CH_UP: ConvTranspose1d(k4, s2, p1)
  int T;                  // tap blocks = weight taps (1 for the im2col'd first conv)
  int pad;                // left padding of the conv (shift of tap i = i - pad)
  int L, log2L;           // GEMM rows per trajectory = INPUT positions of this op
  int in_buf;             // shared-memory activation buffer holding the A operand (0..2), or CH_IN_IM2COL (built from x_t)
  int out_buf;            // 0..2, CH_OUT_GLOBAL (bf16 hi/lo rows to out_hi/out_lo) or CH_OUT_HEAD (fused 1x1 head)
  int res_kind, res_buf;  // what is added after GroupNorm + Mish
  int temb_off;           // >= 0: + temb_rows[b, off + c] + temb2[phase][off + c]
  int phase;              // 0: belongs to the evaluation that ends at the head;  1: to the evaluation that starts after the fused scheduler step
  int gn;                 // GroupNorm(8) + Mish
  int ksteps;             // K steps of 16 channels to issue (4; 3 for the im2col conv when 5*D <= 48)
  uint32_t w_off;         // byte offset of the pre-swizzled weight image in wpack: hi [T*64 rows][128 B], then lo (bf16x3 only)
  const float* bias; const float* gamma; const float* beta;   // [64]
};

struct ChainArgs {
  int n_ops;
  ChainOp ops[CH_MAXOPS];
  int B, H, D;                         // batch rows of this launch; horizon (16) and transition dim of x
  int ns;                              // trajectories per CTA: 8 (full tiles) or 4 (half tiles; measured slower, see api.cu)
  const uint8_t* wpack;
  // first op reading its A operand from global memory (an evaluation's tail: the output of the last per-layer launch)
  const __nv_bfloat16* in_hi; const __nv_bfloat16* in_lo;   // [B, ops[0].L, 64]
  const float* res_f32;                // CH_RES_F32 rows [B * L, 64]
  // x_t (im2col source, 1x1 projection source, scheduler sample)
  const float* x; int x_period;        // row b reads x row (b % x_period) when x_period > 0 (classifier-free guidance feeds [x; x])
  const float* xprojW; const float* xprojB;   // [D][64], [64]
  const float* temb_rows; int temb_stride;
  const float* temb2[2];               // per-step time vectors of the two evaluations a seam launch touches (may be null)
  // fused 1x1 head
  const float* headW; const float* headB; int head_dim; float* head_out;   // [64][head_dim], [head_dim]; head_out [B*H, head_dim] may be null when the step is fused
  // fused scheduler step (after the head; head_dim == D)
  int do_sched;
  SchedK sk;                           // coefficients / flags / noise, traj, mask pointers; mo, sample, n are ignored (taken from shared memory)
  float* x_out;                        // x_{t-1} [B, H, D] (the plan's state buffer, or its output buffer on the last step)
  // evaluation head -> next per-layer launch
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // CH_OUT_GLOBAL rows [B, L/2, 64]
  unsigned long long* trace;           // developer stage clocks of CTA 0: [op][8] (b2p_debug_chain_trace), usually null
};

size_t chain64_smem_bytes();
int launch_chain64(const ChainArgs& a, int nsplit, cudaStream_t s);

}  // namespace b2p
