// EXPERIMENTAL (opt-in, B2P_CLUSTER_EVAL=1; not on any default path): whole-denoiser evaluation of ONE trajectory by ONE
// 16-CTA thread-block cluster.  See unet_cluster.cu for the design.  Internal header (not part of include/b200plan.h).
#pragma once
#include <stdint.h>

#include <cuda_runtime.h>

namespace b2p {

constexpr int UC_CL = 16;              // CTAs per cluster (non-portable cluster size)
constexpr int UC_NT = 512;             // threads per CTA
constexpr int UC_NSTAGE = 4;           // weight ring stages
constexpr int UC_STAGE_FLOATS = 8192;  // 32 KB per stage
constexpr int UC_SLOT_FLOATS = 1024;   // one activation tensor of one trajectory ([L][C], L*C <= 1024)
constexpr int UC_NSLOT = 8;
constexpr int UC_MAXOPS = 48;
constexpr int UC_MAXCHUNKS = 192;
constexpr int UC_MAXL = 16;

// one fused layer (Conv1dBlock / strided conv / transposed conv, + residual, + time term), seen by every CTA of the cluster
struct UcOp {
  int in0, in1, C0, C1;            // input slots (in1 < 0: none) and their channel counts
  int Lin, Lout, Cout, nc;         // nc = Cout / UC_CL output channels per CTA
  int ntaps, jmin, stride, pad, transposed;   // taps [jmin, jmin + ntaps) of the kernel can reach a valid position
  int gn;                          // GroupNorm(8) + Mish on conv + bias
  int temb_off;                    // >= 0: + time-embedding term (column offset), after Mish
  int res_id;                      // >= 0: + identity residual (slot)
  int rin0, rin1, RC0, RC1;        // residual 1x1 conv inputs (RC0 + RC1 == 0: none)
  int out;                         // output slot
  int bias, gamma, beta, resB;     // float offsets into the fp32 pack ([Cout] vectors; < 0: absent)
  int chunk0, nchunks;             // conv weight chunks of this op in UcProgram::chunks (consumed in program order)
  int rnchunks;                    // chunks of the residual 1x1 weights, right after the conv chunks (0: none)
  int head;                        // 1: the 1x1 head follows (final_conv.1), written to global memory
};

// one bulk copy of the per-CTA weight stream: [nc][kstride] floats covering K range [k0, k0 + klen) of the op
struct UcChunk {
  int k0, klen, kstride;
  int off;                         // float offset inside the CTA's stream
  int bytes;                       // nc * kstride * 4
};

struct UcProgram {
  int n_ops, n_chunks;
  int stream_floats_per_cta;       // every CTA's stream has the same layout
  int x_slot;                      // slot the input trajectory is loaded into
  int head_dim, headWk, headB;     // 1x1 head: K-major [head_dim][64] weights and bias (offsets into the pack)
  int pad_;
  UcOp ops[UC_MAXOPS];
  UcChunk chunks[UC_MAXCHUNKS];
};

struct UcLaunch {
  const UcProgram* prog;           // device copy
  const float* stream;             // [UC_CL][stream_floats_per_cta]
  const float* pack;               // fp32 pack (per-channel vectors, head weights)
  const float* x;                  // [B][H][D]
  const float* temb;               // per-sample term  [B][temb_stride]
  int temb_stride;
  const float* temb2;              // per-step term [temb_total] or null
  float* head_out;                 // [B][H][head_dim]
  int B, H, D;
  int dbg;                         // developer timing bisect: 2 skip the dot products, 4 skip the exchange, 8 skip the weight stream; 16: UNTESTED st.async exchange variant
};

size_t uc_smem_bytes();
int launch_unet_cluster(const UcLaunch& a, cudaStream_t s);

}  // namespace b2p
