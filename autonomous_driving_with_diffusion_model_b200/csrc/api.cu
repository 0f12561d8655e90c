// libb200plan: handle, weight packing, denoiser program, whole-plan CUDA graph.  Public ABI: include/b200plan.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "chain64.cuh"

namespace b2p {
int upload_freq_table(const float* f, int n);

enum { BUF_X = -2, BUF_NONE = -1 };
constexpr size_t NPOS = (size_t)-1;

struct Slot {
  std::string key;
  std::vector<int64_t> shape;
  int64_t numel = 0;
  std::vector<float> host;
  bool set = false;
};

struct LayerOp {
  int in0 = BUF_NONE, in1 = BUF_NONE;
  int C0 = 0, C1 = 0, Lin = 0, Lout = 0, Cout = 0, taps = 1, stride = 1, pad = 0, transposed = 0;
  size_t W = NPOS, bias = NPOS, gamma = NPOS, beta = NPOS;
  int temb_off = -1;
  int res_id = BUF_NONE;
  int rin0 = BUF_NONE, rin1 = BUF_NONE, RC0 = 0, RC1 = 0;
  size_t resW = NPOS, resB = NPOS;
  size_t headW = NPOS, headB = NPOS;
  int head_dim = 0;
  int out = BUF_NONE;
  // tensor-core copies (bf16 elements into the 16-bit pack): [taps][Cout][Cin] hi / lo, residual [Cout][RCin] hi / lo
  size_t tcW_hi = NPOS, tcW_lo = NPOS, tcRW_hi = NPOS, tcRW_lo = NPOS;
  // K-major fp32 copies for the small-batch GEMV path: [Cout][taps][Cin], residual [Cout][RCin], head [head_dim][Cout]
  size_t Wk = NPOS, resWk = NPOS, headWk = NPOS;
};

struct Buf { int L, C; size_t off; };  // per-sample floats = L*C; off = prefix sum of per-sample floats

struct GraphKey {
  int B, T, kind, has_target, has_noise, has_traj, has_mask, dev_noise;
  b2p_plan_config pc;
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; int64_t launches; };

}  // namespace b2p

using namespace b2p;

struct b2p_handle_s {
  b2p_model_config cfg;
  int device = 0;
  std::vector<Slot> slots;
  std::unordered_map<std::string, int> index;
  bool finalized = false;
  std::string err;
  int64_t last_launches = 0;
  int64_t flops_per_sample = 0;

  // architecture
  int H = 16, D = 7, dim = 64, nlev = 4;
  int chans[9];
  int temb_total = 0;
  std::vector<LayerOp> ops;
  std::vector<Buf> bufs;
  size_t buf_floats_per_sample = 0;
  int head_dim = 7;

  // packed weights
  std::vector<float> pack_host;
  float* d_pack = nullptr;
  std::vector<uint16_t> pack16_host;   // bf16 hi/lo weight copies for the tcgen05 path
  uint16_t* d_pack16 = nullptr;
  float* d_res0 = nullptr;             // fp32 residual projection of the first block (C_in = transition_dim)
  size_t o_w1t, o_b1, o_w3t, o_b3, o_wc0t, o_bc0, o_wc2t, o_bc2, o_tembW, o_tembB;
  TrajPredWeights tp{};
  bool has_tp = false;

  // row-owned chain kernel (chain64.cu): the 64-channel layers at the full-resolution end of the U-Net
  bool chain_ok = false;
  bool chain_on = true;                // b2p_set_chain / B2P_CHAIN: off = every layer is its own launch
  unsigned long long* d_chain_trace = nullptr;   // developer stage clocks (B2P_CHAIN_TRACE=1)
  size_t l2_window_bytes = 0;          // persisting-L2 window over d_pack16 (0: disabled, B2P_L2_WINDOW=0)
  int chain_u0 = -1;                   // index of the last up level's first conv (runs per-layer, with the residual projection as a sixth tap)
  std::vector<uint32_t> chain_woff;    // per op: byte offset of its pre-swizzled weight image in d_chain (0xffffffff: not a chain op)
  std::vector<uint8_t> chain_host;
  uint8_t* d_chain = nullptr;
  size_t aux_hi = NPOS, aux_lo = NPOS; // 6-tap weights of op chain_u0 in the 16-bit pack: taps 0..4 = conv, tap 5 = residual_conv

  // workspace (activation buffers) for `cap` denoiser rows
  int cap = 0;
  float* d_ws = nullptr;
  float *d_time_embed = nullptr, *d_mish_te = nullptr, *d_mish_feat = nullptr, *d_temb = nullptr, *d_act = nullptr;
  int64_t* d_t = nullptr;

  // static plan buffers
  int plan_capB = 0, plan_capT = 0;
  float *p_x = nullptr, *p_feat = nullptr, *p_target = nullptr, *p_cond = nullptr, *p_noise = nullptr, *p_traj = nullptr,
        *p_mask = nullptr, *p_mo = nullptr, *p_action = nullptr, *p_out = nullptr, *p_ttab = nullptr, *p_itab = nullptr, *p_tetab = nullptr;
  int64_t* p_tsteps = nullptr;
  unsigned long long* p_seed = nullptr;   // Philox key of the current plan (device; read by the scheduler kernels at run time)
  float* p_thr = nullptr;                 // dynamic-threshold scratch [B]
  unsigned long long noise_seed = 0x5eed5eed5eedULL, noise_calls = 0, last_noise_key = 0;
  int ts_ntrain = -1, ts_T = -1;       // which timestep table p_tsteps currently holds (the captured kernels read it at replay time)
  std::vector<GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  int small_batch_max = B2P_SMALL_BATCH_DEFAULT;   // largest evaluation batch that takes the GEMV kernels

  int fail(int code, const std::string& m) { err = m; return code; }
};

namespace {

struct Packer {
  std::vector<float>& v;
  size_t push(const float* p, size_t n) {
    size_t off = (v.size() + 63) & ~(size_t)63;
    v.resize(off + n);
    if (p) memcpy(v.data() + off, p, n * sizeof(float));
    return off;
  }
  size_t alloc(size_t n) { return push(nullptr, n); }
};

void add_slot(b2p_handle_s* h, const std::string& key, std::vector<int64_t> shape) {
  Slot s;
  s.key = key; s.shape = shape; s.numel = 1;
  for (auto d : shape) s.numel *= d;
  h->index[key] = (int)h->slots.size();
  h->slots.push_back(std::move(s));
}
void add_conv_block_slots(b2p_handle_s* h, const std::string& p, int cin, int cout, int k) {
  add_slot(h, p + ".block.0.weight", {cout, cin, k});
  add_slot(h, p + ".block.0.bias", {cout});
  add_slot(h, p + ".block.2.weight", {cout});
  add_slot(h, p + ".block.2.bias", {cout});
}
void add_res_block_slots(b2p_handle_s* h, const std::string& p, int cin, int cout) {
  add_conv_block_slots(h, p + ".blocks.0", cin, cout, 5);
  add_conv_block_slots(h, p + ".blocks.1", cout, cout, 5);
  add_slot(h, p + ".time_mlp.1.weight", {cout, 2 * h->dim});
  add_slot(h, p + ".time_mlp.1.bias", {cout});
  if (cin != cout) {
    add_slot(h, p + ".residual_conv.weight", {cout, cin, 1});
    add_slot(h, p + ".residual_conv.bias", {cout});
  }
}

const float* W(b2p_handle_s* h, const std::string& key) { return h->slots[h->index.at(key)].host.data(); }

// Conv1d weight [Cout][Cin][k] -> [k][Cin][Cout]
size_t pack_conv(b2p_handle_s* h, Packer& pk, const std::string& key, int cout, int cin, int k) {
  const float* w = W(h, key);
  size_t off = pk.alloc((size_t)k * cin * cout);
  float* o = pk.v.data() + off;
  for (int co = 0; co < cout; ++co)
    for (int c = 0; c < cin; ++c)
      for (int j = 0; j < k; ++j) o[((size_t)j * cin + c) * cout + co] = w[((size_t)co * cin + c) * k + j];
  return off;
}
// ConvTranspose1d weight [Cin][Cout][k] -> [k][Cin][Cout]
size_t pack_convT(b2p_handle_s* h, Packer& pk, const std::string& key, int cin, int cout, int k) {
  const float* w = W(h, key);
  size_t off = pk.alloc((size_t)k * cin * cout);
  float* o = pk.v.data() + off;
  for (int c = 0; c < cin; ++c)
    for (int co = 0; co < cout; ++co)
      for (int j = 0; j < k; ++j) o[((size_t)j * cin + c) * cout + co] = w[((size_t)c * cout + co) * k + j];
  return off;
}
// Linear weight [out][in] -> [in][out_ld] placed at column col0 of a wider matrix
void pack_linear_T(const float* w, int out, int in, float* dst, int ld, int col0) {
  for (int o = 0; o < out; ++o)
    for (int i = 0; i < in; ++i) dst[(size_t)i * ld + col0 + o] = w[(size_t)o * in + i];
}
size_t pack_vec(b2p_handle_s* h, Packer& pk, const std::string& key, size_t n) { return pk.push(W(h, key), n); }

// Conv1d weight [Cout][Cin][k] (or ConvTranspose1d [Cin][Cout][k]) -> K-major fp32 [Cout][k][Cin] (GEMV path)
size_t pack_conv_kmajor(b2p_handle_s* h, Packer& pk, const std::string& key, int cout, int cin, int k, bool transposed) {
  const float* w = W(h, key);
  size_t off = pk.alloc((size_t)k * cin * cout);
  float* o = pk.v.data() + off;
  for (int co = 0; co < cout; ++co)
    for (int j = 0; j < k; ++j)
      for (int c = 0; c < cin; ++c)
        o[((size_t)co * k + j) * cin + c] = transposed ? w[((size_t)c * cout + co) * k + j] : w[((size_t)co * cin + c) * k + j];
  return off;
}

// ---- bf16 hi/lo split (round-to-nearest-even), host side ----
inline uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
size_t alloc16(std::vector<uint16_t>& v, size_t n) {
  size_t off = (v.size() + 127) & ~(size_t)127;   // 256-byte aligned (TMA global address alignment)
  v.resize(off + n);
  return off;
}
// Conv1d weight [Cout][Cin][k] (or ConvTranspose1d [Cin][Cout][k]) -> K-major [k][Cout][Cin] bf16 hi and lo
void pack_conv_tc(b2p_handle_s* h, const std::string& key, int cout, int cin, int k, bool transposed, size_t* hi, size_t* lo) {
  const float* w = W(h, key);
  size_t n = (size_t)k * cout * cin;
  *hi = alloc16(h->pack16_host, n);
  *lo = alloc16(h->pack16_host, n);
  uint16_t* ph = h->pack16_host.data() + *hi;
  uint16_t* pl = h->pack16_host.data() + *lo;
  for (int j = 0; j < k; ++j)
    for (int co = 0; co < cout; ++co)
      for (int c = 0; c < cin; ++c) {
        float v = transposed ? w[((size_t)c * cout + co) * k + j] : w[((size_t)co * cin + c) * k + j];
        size_t o = ((size_t)j * cout + co) * cin + c;
        ph[o] = f2bf(v);
        pl[o] = f2bf(v - bf2f(ph[o]));
      }
}

int new_buf(b2p_handle_s* h, int L, int C) {
  Buf b{L, C, h->buf_floats_per_sample};
  h->buf_floats_per_sample += (size_t)L * C;
  h->bufs.push_back(b);
  return (int)h->bufs.size() - 1;
}

int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

// ---- build the slot table (state_dict contract) ----
void build_slots(b2p_handle_s* h) {
  const int dim = h->dim;
  if (h->cfg.guidance == B2P_FREE_GUIDANCE) {
    add_slot(h, "cond_mlp.0.weight", {dim, 2}); add_slot(h, "cond_mlp.0.bias", {dim});
    add_slot(h, "cond_mlp.2.weight", {dim, dim}); add_slot(h, "cond_mlp.2.bias", {dim});
  }
  add_slot(h, "time_mlp.1.weight", {4 * dim, dim}); add_slot(h, "time_mlp.1.bias", {4 * dim});
  add_slot(h, "time_mlp.3.weight", {dim, 4 * dim}); add_slot(h, "time_mlp.3.bias", {dim});
  const int n = h->nlev;
  for (int i = 0; i < n; ++i) {
    int ci = h->chans[i], co = h->chans[i + 1];
    std::string p = "downs." + std::to_string(i);
    add_res_block_slots(h, p + ".0", ci, co);
    add_res_block_slots(h, p + ".1", co, co);
    if (i < n - 1) { add_slot(h, p + ".3.conv.weight", {co, co, 3}); add_slot(h, p + ".3.conv.bias", {co}); }
  }
  for (int u = 0; u < n - 1; ++u) {
    int ci = h->chans[n - 1 - u], co = h->chans[n - u];  // (dim_in, dim_out) of reversed(in_out[1:])
    std::string p = "ups." + std::to_string(u);
    add_res_block_slots(h, p + ".0", co * 2, ci);
    add_res_block_slots(h, p + ".1", ci, ci);
    add_slot(h, p + ".3.conv.weight", {ci, ci, 4}); add_slot(h, p + ".3.conv.bias", {ci});
  }
  int mid = h->chans[n];
  add_res_block_slots(h, "mid_block1", mid, mid);
  add_res_block_slots(h, "mid_block2", mid, mid);
  int fin = h->chans[1];
  if (h->cfg.guidance == B2P_CLASSIFIER_GUIDANCE) {
    add_conv_block_slots(h, "act_conv.0", fin, fin, 5);
    add_slot(h, "act_conv.1.weight", {3, fin, 1}); add_slot(h, "act_conv.1.bias", {3});
    const int hd = 64;
    add_slot(h, "state_pred.input_proj.weight", {hd, 3}); add_slot(h, "state_pred.input_proj.bias", {hd});
    for (int l = 0; l < 2; ++l) {
      std::string p = "state_pred.encoder_traj.layers." + std::to_string(l);
      add_slot(h, p + ".self_attn.in_proj_weight", {3 * hd, hd}); add_slot(h, p + ".self_attn.in_proj_bias", {3 * hd});
      add_slot(h, p + ".self_attn.out_proj.weight", {hd, hd}); add_slot(h, p + ".self_attn.out_proj.bias", {hd});
      add_slot(h, p + ".linear1.weight", {4 * hd, hd}); add_slot(h, p + ".linear1.bias", {4 * hd});
      add_slot(h, p + ".linear2.weight", {hd, 4 * hd}); add_slot(h, p + ".linear2.bias", {hd});
      add_slot(h, p + ".norm1.weight", {hd}); add_slot(h, p + ".norm1.bias", {hd});
      add_slot(h, p + ".norm2.weight", {hd}); add_slot(h, p + ".norm2.bias", {hd});
    }
    add_slot(h, "state_pred.encoder_traj.norm.weight", {hd}); add_slot(h, "state_pred.encoder_traj.norm.bias", {hd});
    add_slot(h, "state_pred.output_proj.weight", {h->D - 3, hd}); add_slot(h, "state_pred.output_proj.bias", {h->D - 3});
  } else {
    add_conv_block_slots(h, "final_conv.0", fin, fin, 5);
    add_slot(h, "final_conv.1.weight", {h->D, fin, 1}); add_slot(h, "final_conv.1.bias", {h->D});
  }
}

// ---- residual block = two fused launches ----
struct BlockBuild { int temb_off; };
int add_res_block(b2p_handle_s* h, Packer& pk, const std::string& p, int in0, int in1, int C0, int C1, int cout, int L,
                  int& temb_cursor, std::vector<float>& tembW, std::vector<float>& tembB) {
  const int cin = C0 + C1;
  LayerOp a;
  a.in0 = in0; a.in1 = in1; a.C0 = C0; a.C1 = C1; a.Lin = a.Lout = L; a.Cout = cout; a.taps = 5; a.stride = 1; a.pad = 2;
  a.W = pack_conv(h, pk, p + ".blocks.0.block.0.weight", cout, cin, 5);
  a.Wk = pack_conv_kmajor(h, pk, p + ".blocks.0.block.0.weight", cout, cin, 5, false);
  if (cin % 64 == 0) pack_conv_tc(h, p + ".blocks.0.block.0.weight", cout, cin, 5, false, &a.tcW_hi, &a.tcW_lo);
  a.bias = pack_vec(h, pk, p + ".blocks.0.block.0.bias", cout);
  a.gamma = pack_vec(h, pk, p + ".blocks.0.block.2.weight", cout);
  a.beta = pack_vec(h, pk, p + ".blocks.0.block.2.bias", cout);
  a.temb_off = temb_cursor;
  // this block's Linear(2*dim -> cout) becomes columns [temb_cursor, temb_cursor+cout) of the shared GEMM
  pack_linear_T(W(h, p + ".time_mlp.1.weight"), cout, 2 * h->dim, tembW.data(), h->temb_total, temb_cursor);
  memcpy(tembB.data() + temb_cursor, W(h, p + ".time_mlp.1.bias"), sizeof(float) * cout);
  temb_cursor += cout;
  a.out = new_buf(h, L, cout);
  h->ops.push_back(a);

  LayerOp b;
  b.in0 = a.out; b.C0 = cout; b.Lin = b.Lout = L; b.Cout = cout; b.taps = 5; b.stride = 1; b.pad = 2;
  b.W = pack_conv(h, pk, p + ".blocks.1.block.0.weight", cout, cout, 5);
  b.Wk = pack_conv_kmajor(h, pk, p + ".blocks.1.block.0.weight", cout, cout, 5, false);
  pack_conv_tc(h, p + ".blocks.1.block.0.weight", cout, cout, 5, false, &b.tcW_hi, &b.tcW_lo);
  b.bias = pack_vec(h, pk, p + ".blocks.1.block.0.bias", cout);
  b.gamma = pack_vec(h, pk, p + ".blocks.1.block.2.weight", cout);
  b.beta = pack_vec(h, pk, p + ".blocks.1.block.2.bias", cout);
  if (cin != cout) {
    b.rin0 = in0; b.rin1 = in1; b.RC0 = C0; b.RC1 = C1;
    b.resW = pack_conv(h, pk, p + ".residual_conv.weight", cout, cin, 1);
    b.resWk = pack_conv_kmajor(h, pk, p + ".residual_conv.weight", cout, cin, 1, false);
    if (cin % 64 == 0) pack_conv_tc(h, p + ".residual_conv.weight", cout, cin, 1, false, &b.tcRW_hi, &b.tcRW_lo);
    b.resB = pack_vec(h, pk, p + ".residual_conv.bias", cout);
  } else {
    b.res_id = in0;
  }
  b.out = new_buf(h, L, cout);
  h->ops.push_back(b);
  // FLOPs: two k=5 convs + time Linear + optional 1x1
  h->flops_per_sample += 2LL * cout * cin * 5 * L + 2LL * cout * cout * 5 * L + 2LL * cout * 2 * h->dim;
  if (cin != cout) h->flops_per_sample += 2LL * cout * cin * L;
  return b.out;
}

int build_program(b2p_handle_s* h) {
  h->ops.clear(); h->bufs.clear(); h->buf_floats_per_sample = 0; h->flops_per_sample = 0;
  h->pack_host.clear(); h->pack16_host.clear();
  Packer pk{h->pack_host};
  const int dim = h->dim, n = h->nlev;
  // embedding MLPs, transposed for coalesced reads
  {
    std::vector<float> t((size_t)dim * 4 * dim);
    pack_linear_T(W(h, "time_mlp.1.weight"), 4 * dim, dim, t.data(), 4 * dim, 0);
    h->o_w1t = pk.push(t.data(), t.size());
    h->o_b1 = pack_vec(h, pk, "time_mlp.1.bias", 4 * dim);
    pack_linear_T(W(h, "time_mlp.3.weight"), dim, 4 * dim, t.data(), dim, 0);
    h->o_w3t = pk.push(t.data(), t.size());
    h->o_b3 = pack_vec(h, pk, "time_mlp.3.bias", dim);
    h->flops_per_sample += 2LL * dim * 4 * dim * 2;
    if (h->cfg.guidance == B2P_FREE_GUIDANCE) {
      std::vector<float> c((size_t)dim * dim);
      pack_linear_T(W(h, "cond_mlp.0.weight"), dim, 2, c.data(), dim, 0);
      h->o_wc0t = pk.push(c.data(), 2 * dim);
      h->o_bc0 = pack_vec(h, pk, "cond_mlp.0.bias", dim);
      pack_linear_T(W(h, "cond_mlp.2.weight"), dim, dim, c.data(), dim, 0);
      h->o_wc2t = pk.push(c.data(), (size_t)dim * dim);
      h->o_bc2 = pack_vec(h, pk, "cond_mlp.2.bias", dim);
      h->flops_per_sample += 2LL * dim * 2 + 2LL * dim * dim;
    }
  }
  std::vector<float> tembW((size_t)2 * dim * h->temb_total), tembB(h->temb_total);
  int cursor = 0;
  int L = h->H;
  int x = BUF_X, xc = h->D;
  std::vector<int> skips, skipC, skipL;
  for (int i = 0; i < n; ++i) {
    int co = h->chans[i + 1];
    std::string p = "downs." + std::to_string(i);
    x = add_res_block(h, pk, p + ".0", x, BUF_NONE, xc, 0, co, L, cursor, tembW, tembB);
    x = add_res_block(h, pk, p + ".1", x, BUF_NONE, co, 0, co, L, cursor, tembW, tembB);
    xc = co;
    skips.push_back(x); skipC.push_back(co); skipL.push_back(L);
    if (i < n - 1) {
      LayerOp d;
      d.in0 = x; d.C0 = co; d.Lin = L; d.Lout = L / 2; d.Cout = co; d.taps = 3; d.stride = 2; d.pad = 1;
      d.W = pack_conv(h, pk, p + ".3.conv.weight", co, co, 3);
      d.Wk = pack_conv_kmajor(h, pk, p + ".3.conv.weight", co, co, 3, false);
      pack_conv_tc(h, p + ".3.conv.weight", co, co, 3, false, &d.tcW_hi, &d.tcW_lo);
      d.bias = pack_vec(h, pk, p + ".3.conv.bias", co);
      d.out = new_buf(h, L / 2, co);
      h->ops.push_back(d);
      h->flops_per_sample += 2LL * co * co * 3 * (L / 2);
      x = d.out; L /= 2;
    }
  }
  x = add_res_block(h, pk, "mid_block1", x, BUF_NONE, xc, 0, xc, L, cursor, tembW, tembB);
  x = add_res_block(h, pk, "mid_block2", x, BUF_NONE, xc, 0, xc, L, cursor, tembW, tembB);
  for (int u = 0; u < n - 1; ++u) {
    int ci = h->chans[n - 1 - u];
    std::string p = "ups." + std::to_string(u);
    int sk = skips.back(), sc = skipC.back();
    skips.pop_back(); skipC.pop_back(); skipL.pop_back();
    x = add_res_block(h, pk, p + ".0", x, sk, xc, sc, ci, L, cursor, tembW, tembB);  // cat((x, skip), dim=1)
    x = add_res_block(h, pk, p + ".1", x, BUF_NONE, ci, 0, ci, L, cursor, tembW, tembB);
    xc = ci;
    LayerOp t;
    t.in0 = x; t.C0 = ci; t.Lin = L; t.Lout = 2 * L; t.Cout = ci; t.taps = 4; t.stride = 2; t.pad = 1; t.transposed = 1;
    t.W = pack_convT(h, pk, p + ".3.conv.weight", ci, ci, 4);
    t.Wk = pack_conv_kmajor(h, pk, p + ".3.conv.weight", ci, ci, 4, true);
    pack_conv_tc(h, p + ".3.conv.weight", ci, ci, 4, true, &t.tcW_hi, &t.tcW_lo);
    t.bias = pack_vec(h, pk, p + ".3.conv.bias", ci);
    t.out = new_buf(h, 2 * L, ci);
    h->ops.push_back(t);
    h->flops_per_sample += 2LL * ci * ci * 4 * L;
    x = t.out; L *= 2;
  }
  if (cursor != h->temb_total) return B2P_ERR_INVALID_ARG;
  {  // head: Conv1dBlock(fin, fin, 5) + 1x1 conv, fused
    const bool cls = h->cfg.guidance == B2P_CLASSIFIER_GUIDANCE;
    std::string p = cls ? "act_conv" : "final_conv";
    int fin = h->chans[1];
    if (fin != 64) return B2P_ERR_INVALID_ARG;  // fused head needs the whole channel vector in one tile
    LayerOp f;
    f.in0 = x; f.C0 = fin; f.Lin = f.Lout = L; f.Cout = fin; f.taps = 5; f.stride = 1; f.pad = 2;
    f.W = pack_conv(h, pk, p + ".0.block.0.weight", fin, fin, 5);
    f.Wk = pack_conv_kmajor(h, pk, p + ".0.block.0.weight", fin, fin, 5, false);
    f.headWk = pack_vec(h, pk, p + ".1.weight", (size_t)(cls ? 3 : h->D) * fin);   // Conv1d(fin, head_dim, 1) weight is already [head_dim][fin]
    f.out = new_buf(h, L, fin);   // output of the conv block (only the GEMV path, which runs the 1x1 head as its own launch, reads it)
    pack_conv_tc(h, p + ".0.block.0.weight", fin, fin, 5, false, &f.tcW_hi, &f.tcW_lo);
    f.bias = pack_vec(h, pk, p + ".0.block.0.bias", fin);
    f.gamma = pack_vec(h, pk, p + ".0.block.2.weight", fin);
    f.beta = pack_vec(h, pk, p + ".0.block.2.bias", fin);
    h->head_dim = cls ? 3 : h->D;
    std::vector<float> hw((size_t)fin * h->head_dim);
    pack_linear_T(W(h, p + ".1.weight"), h->head_dim, fin, hw.data(), h->head_dim, 0);
    f.headW = pk.push(hw.data(), hw.size());
    f.headB = pack_vec(h, pk, p + ".1.bias", h->head_dim);
    f.head_dim = h->head_dim;
    h->ops.push_back(f);
    h->flops_per_sample += 2LL * fin * fin * 5 * L + 2LL * h->head_dim * fin * L;
  }
  h->o_tembW = pk.push(tembW.data(), tembW.size());
  h->o_tembB = pk.push(tembB.data(), tembB.size());
  return B2P_OK;
}

// ---- chain kernel: which ops it covers and their pre-swizzled weight images ----
// image of one op: plane hi = [T*64 rows (channel quarter, tap, 16 out channels)][64 k] bf16, K-major with the 128-byte swizzle applied (what TMA would
// have produced in shared memory), followed by plane lo.  The kernel copies it with plain bulk copies, one per tap.
uint32_t chain_add_image(b2p_handle_s* h, const float* w, int T, int cin, int ktaps, bool transposed, int im2col_D) {
  const size_t plane = (size_t)T * 64 * 128;
  const size_t off = (h->chain_host.size() + 1023) & ~(size_t)1023;
  h->chain_host.resize(off + 2 * plane, 0);
  uint8_t* base = h->chain_host.data() + off;
  for (int t = 0; t < T; ++t)
    for (int co = 0; co < 64; ++co)
      for (int k = 0; k < 64; ++k) {
        float v = 0.f;
        if (im2col_D > 0) {            // first conv: k = j * D + c over the 5 taps of Conv1d [64][D][5]
          const int j = k / im2col_D, c = k - j * im2col_D;
          if (j < ktaps) v = w[((size_t)co * im2col_D + c) * ktaps + j];
        } else if (k < cin) {
          v = transposed ? w[((size_t)k * 64 + co) * ktaps + t] : w[((size_t)co * cin + k) * ktaps + t];
        }
        const int n = (co >> 4) * (T * 16) + t * 16 + (co & 15);   // [channel quarter][tap][16 channels]: the taps of one quarter (or of two adjacent quarters = a half) are one MMA operand
        const size_t o = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
        const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
        memcpy(base + o, &hi, 2);
        memcpy(base + plane + o, &lo, 2);
      }
  return (uint32_t)off;
}

void build_chain(b2p_handle_s* h) {
  h->chain_ok = false; h->chain_u0 = -1; h->chain_host.clear(); h->aux_hi = h->aux_lo = NPOS;
  h->chain_woff.assign(h->ops.size(), 0xffffffffu);
  const int n = h->nlev, nops = (int)h->ops.size();
  if (h->H != 16 || h->chans[1] != 64 || h->D > 8 || n < 2) return;
  const int u0 = nops - 6;   // [.., ups.last.0.conv0, .conv1, ups.last.1.conv0, .conv1, Upsample1d, head]
  if (u0 < 5) return;
  auto is64 = [&](const LayerOp& op, int L, int taps) { return op.Cout == 64 && op.Lin == L && op.taps == taps; };
  const LayerOp* o = h->ops.data();
  if (!(is64(o[0], 16, 5) && o[0].C0 == h->D && is64(o[1], 16, 5) && o[1].resW != NPOS && o[1].RC0 == h->D && is64(o[2], 16, 5) && is64(o[3], 16, 5) &&
        o[3].res_id == o[1].out && is64(o[4], 16, 3) && o[4].stride == 2)) return;
  if (!(is64(o[u0], 8, 5) && o[u0].C1 > 0 && is64(o[u0 + 1], 8, 5) && o[u0 + 1].resW != NPOS && o[u0 + 1].tcRW_hi != NPOS && is64(o[u0 + 2], 8, 5) &&
        is64(o[u0 + 3], 8, 5) && o[u0 + 3].res_id == o[u0 + 1].out && is64(o[u0 + 4], 8, 4) && o[u0 + 4].transposed && is64(o[u0 + 5], 16, 5) &&
        o[u0 + 5].headW != NPOS)) return;
  const std::string up = "ups." + std::to_string(n - 2), head = h->cfg.guidance == B2P_CLASSIFIER_GUIDANCE ? "act_conv" : "final_conv";
  const char* keyA[5] = {"downs.0.0.blocks.0.block.0.weight", "downs.0.0.blocks.1.block.0.weight", "downs.0.1.blocks.0.block.0.weight",
                         "downs.0.1.blocks.1.block.0.weight", "downs.0.3.conv.weight"};
  h->chain_woff[0] = chain_add_image(h, W(h, keyA[0]), 1, h->D, 5, false, h->D);
  for (int i = 1; i < 4; ++i) h->chain_woff[i] = chain_add_image(h, W(h, keyA[i]), 5, 64, 5, false, 0);
  h->chain_woff[4] = chain_add_image(h, W(h, keyA[4]), 3, 64, 3, false, 0);
  h->chain_woff[u0 + 1] = chain_add_image(h, W(h, up + ".0.blocks.1.block.0.weight"), 5, 64, 5, false, 0);
  h->chain_woff[u0 + 2] = chain_add_image(h, W(h, up + ".1.blocks.0.block.0.weight"), 5, 64, 5, false, 0);
  h->chain_woff[u0 + 3] = chain_add_image(h, W(h, up + ".1.blocks.1.block.0.weight"), 5, 64, 5, false, 0);
  h->chain_woff[u0 + 4] = chain_add_image(h, W(h, up + ".3.conv.weight"), 4, 64, 4, true, 0);
  h->chain_woff[u0 + 5] = chain_add_image(h, W(h, head + ".0.block.0.weight"), 5, 64, 5, false, 0);
  {  // op u0 with the block's residual projection as a sixth tap: [6][64][Cin] hi / lo (K-major, read through TMA)
    const int cin = o[u0].C0 + o[u0].C1;
    const float* w = W(h, up + ".0.blocks.0.block.0.weight");
    const float* rw = W(h, up + ".0.residual_conv.weight");
    const size_t nel = (size_t)6 * 64 * cin;
    h->aux_hi = alloc16(h->pack16_host, nel);
    h->aux_lo = alloc16(h->pack16_host, nel);
    uint16_t* ph = h->pack16_host.data() + h->aux_hi;
    uint16_t* pl = h->pack16_host.data() + h->aux_lo;
    for (int j = 0; j < 6; ++j)
      for (int co = 0; co < 64; ++co)
        for (int c = 0; c < cin; ++c) {
          const float v = j < 5 ? w[((size_t)co * cin + c) * 5 + j] : rw[(size_t)co * cin + c];
          const size_t q = ((size_t)j * 64 + co) * cin + c;
          ph[q] = f2bf(v);
          pl[q] = f2bf(v - bf2f(ph[q]));
        }
  }
  h->chain_u0 = u0;
  h->chain_ok = true;
}

void build_trajpred(b2p_handle_s* h, Packer& pk, std::vector<size_t>& offs) {
  // offsets recorded in order; resolved to device pointers after upload
  const int hd = 64, S = h->H - 1;
  auto lin_t = [&](const std::string& key, int out, int in) {
    std::vector<float> t((size_t)in * out);
    pack_linear_T(W(h, key), out, in, t.data(), out, 0);
    return pk.push(t.data(), t.size());
  };
  offs.push_back(lin_t("state_pred.input_proj.weight", hd, 3));
  offs.push_back(pack_vec(h, pk, "state_pred.input_proj.bias", hd));
  {  // positional table: SinusoidalPosEmb(64)(arange(S)) (modeling/helpers.py:54)
    std::vector<float> pos((size_t)S * hd);
    const int half = hd / 2;
    float sc = (float)(-(log(10000.0) / (half - 1)));
    for (int s = 0; s < S; ++s)
      for (int i = 0; i < half; ++i) {
        float f = expf((float)i * sc);
        float arg = (float)s * f;
        pos[(size_t)s * hd + i] = sinf(arg);
        pos[(size_t)s * hd + half + i] = cosf(arg);
      }
    offs.push_back(pk.push(pos.data(), pos.size()));
  }
  for (int l = 0; l < 2; ++l) {
    std::string p = "state_pred.encoder_traj.layers." + std::to_string(l);
    offs.push_back(lin_t(p + ".self_attn.in_proj_weight", 3 * hd, hd));
    offs.push_back(pack_vec(h, pk, p + ".self_attn.in_proj_bias", 3 * hd));
    offs.push_back(lin_t(p + ".self_attn.out_proj.weight", hd, hd));
    offs.push_back(pack_vec(h, pk, p + ".self_attn.out_proj.bias", hd));
    offs.push_back(lin_t(p + ".linear1.weight", 4 * hd, hd));
    offs.push_back(pack_vec(h, pk, p + ".linear1.bias", 4 * hd));
    offs.push_back(lin_t(p + ".linear2.weight", hd, 4 * hd));
    offs.push_back(pack_vec(h, pk, p + ".linear2.bias", hd));
    offs.push_back(pack_vec(h, pk, p + ".norm1.weight", hd));
    offs.push_back(pack_vec(h, pk, p + ".norm1.bias", hd));
    offs.push_back(pack_vec(h, pk, p + ".norm2.weight", hd));
    offs.push_back(pack_vec(h, pk, p + ".norm2.bias", hd));
    offs.push_back(pack_vec(h, pk, p + ".self_attn.in_proj_weight", (size_t)3 * hd * hd));
    offs.push_back(pack_vec(h, pk, p + ".self_attn.out_proj.weight", (size_t)hd * hd));
    offs.push_back(pack_vec(h, pk, p + ".linear1.weight", (size_t)4 * hd * hd));
    offs.push_back(pack_vec(h, pk, p + ".linear2.weight", (size_t)4 * hd * hd));
  }
  offs.push_back(pack_vec(h, pk, "state_pred.encoder_traj.norm.weight", hd));
  offs.push_back(pack_vec(h, pk, "state_pred.encoder_traj.norm.bias", hd));
  offs.push_back(lin_t("state_pred.output_proj.weight", h->D - 3, hd));
  offs.push_back(pack_vec(h, pk, "state_pred.output_proj.bias", h->D - 3));
  offs.push_back(pack_vec(h, pk, "state_pred.output_proj.weight", (size_t)(h->D - 3) * hd));
  offs.push_back(pack_vec(h, pk, "state_pred.input_proj.weight", (size_t)hd * 3));
}

void resolve_trajpred(b2p_handle_s* h, const std::vector<size_t>& o) {
  const float* b = h->d_pack;
  size_t i = 0;
  TrajPredWeights& w = h->tp;
  w.in_w = b + o[i++]; w.in_b = b + o[i++]; w.pos = b + o[i++];
  w.n_layers = 2;
  for (int l = 0; l < 2; ++l) {
    auto& y = w.layer[l];
    y.qkv_wt = b + o[i++]; y.qkv_b = b + o[i++]; y.out_wt = b + o[i++]; y.out_b = b + o[i++];
    y.l1_wt = b + o[i++]; y.l1_b = b + o[i++]; y.l2_wt = b + o[i++]; y.l2_b = b + o[i++];
    y.n1_g = b + o[i++]; y.n1_b = b + o[i++]; y.n2_g = b + o[i++]; y.n2_b = b + o[i++];
    y.qkv_w = b + o[i++]; y.out_w = b + o[i++]; y.l1_w = b + o[i++]; y.l2_w = b + o[i++];
  }
  w.fn_g = b + o[i++]; w.fn_b = b + o[i++]; w.out_wt = b + o[i++]; w.out_b = b + o[i++]; w.out_w = b + o[i++];
  w.in_w_raw = b + o[i++];
  h->has_tp = true;
}

void drop_graphs(b2p_handle_s* h) {
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

int ensure_workspace(b2p_handle_s* h, int rows) {
  if (rows <= h->cap) return B2P_OK;
  drop_graphs(h);
  if (h->d_ws) { B2P_CUDA_TRY(cudaFree(h->d_ws)); h->d_ws = nullptr; }
  int cap = rows;
  size_t per = (size_t)h->dim + 2 * h->dim + h->temb_total + h->buf_floats_per_sample + (size_t)h->H * h->chans[1];
  size_t floats = per * cap + 64 * 16;
  B2P_CUDA_TRY(cudaMalloc((void**)&h->d_ws, floats * sizeof(float) + sizeof(int64_t) * (cap + 8)));
  float* p = h->d_ws;
  auto take = [&](size_t n) { float* r = p; p += (n + 63) & ~(size_t)63; return r; };
  h->d_time_embed = take((size_t)cap * h->dim);
  h->d_mish_te = take((size_t)cap * h->dim);
  h->d_mish_feat = take((size_t)cap * h->dim);
  h->d_temb = take((size_t)cap * h->temb_total);
  h->d_act = take((size_t)cap * h->buf_floats_per_sample);
  h->d_res0 = take((size_t)cap * h->H * h->chans[1]);
  h->d_t = reinterpret_cast<int64_t*>(h->d_ws + floats);
  h->cap = cap;
  return B2P_OK;
}

inline const float* buf_ptr(b2p_handle_s* h, int id, const float* x) {
  if (id == BUF_X) return x;
  if (id == BUF_NONE) return nullptr;
  return h->d_act + h->bufs[id].off * (size_t)h->cap;
}

// the denoiser on `rows` batch rows; x rows may repeat with period x_period (CFG feeds [x; x]).
// itab/ttab_row != null: "table mode" (inside a plan, no CFG): the per-block time-MLP outputs were precomputed as an
// image term itab[rows, temb_total] (step-invariant) and a time vector ttab_row[temb_total] for this step, so the
// embedding kernels are skipped.
// Seam of two evaluations inside a plan without guidance (chain64.cu): behind the tail chain of this evaluation run the scheduler
// step and, unless this is the last step, the head chain of the NEXT evaluation (whose per-step time vector is next_ttab_row).
struct ChainSeam {
  SchedLaunch sched;            // mo / sample are taken from shared memory; prev = where x_{t-1} goes
  bool next_head;
  const float* next_ttab_row;
};

bool chain_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2P_CHAIN"); on = e ? atoi(e) : 1; }
  return on != 0;
}

int run_unet(b2p_handle_s* h, const float* x, int x_period, const float* feat, int feat_rows, const int64_t* t, int t_count,
             const float* cond, float* head_out, float* time_embed_out, int rows, cudaStream_t s, int64_t* launches,
             const float* itab = nullptr, const float* ttab_row = nullptr, const ChainSeam* seam = nullptr, bool head_chain_done = false) {
  if (!h->finalized) return h->fail(B2P_ERR_NOT_FINALIZED, "weights not finalized");
  int rc = ensure_workspace(h, rows);
  if (rc) return rc;
  const float* P = h->d_pack;
  const float* temb_rows = itab;          // per-sample term [rows, temb_total]
  if (!itab) {
    EmbedArgs e{};
    e.t = t; e.t_count = t_count; e.feat = feat; e.feat_rows = feat_rows; e.cond = cond;
    e.use_cond = h->cfg.guidance == B2P_FREE_GUIDANCE;
    e.w1t = P + h->o_w1t; e.b1 = P + h->o_b1; e.w3t = P + h->o_w3t; e.b3 = P + h->o_b3;
    if (e.use_cond) { e.wc0t = P + h->o_wc0t; e.bc0 = P + h->o_bc0; e.wc2t = P + h->o_wc2t; e.bc2 = P + h->o_bc2; }
    e.time_embed = time_embed_out ? time_embed_out : h->d_time_embed;
    e.mish_te = h->d_mish_te; e.mish_feat = h->d_mish_feat; e.te_rows = rows; e.feat_out_rows = rows; e.B = rows; e.dim = h->dim;
    if ((rc = launch_embed(e, s))) return rc;
    ++*launches;
    // all 16 block time-MLPs as one GEMM [rows, 2*dim] x [2*dim, temb_total] (the two halves are two K-slabs)
    ConvArgs a{};
    a.x0 = h->d_mish_te; a.C0 = h->dim; a.x1 = h->d_mish_feat; a.C1 = h->dim;
    a.Lin = a.Lout = 1; a.log2Lout = 0; a.nrows = rows; a.Cout = h->temb_total;
    a.taps = 1; a.jmin = a.jmax = 0; a.stride = 1; a.W = P + h->o_tembW; a.bias = P + h->o_tembB; a.out = h->d_temb;
    if ((rc = launch_conv_ffma(a, s))) return rc;
    ++*launches;
    temb_rows = h->d_temb;
  }
  // Small batches take the exact-fp32 GEMV program whatever the precision mode: with a handful of trajectories a layer is
  // pure weight streaming, the GEMV kernels prefetch weights layers ahead, and no bf16 splitting is needed.
  const bool use_gemv = rows <= h->small_batch_max;
  const int prec = use_gemv ? (int)B2P_PREC_FP32 : h->cfg.precision;
  const bool tc = prec != B2P_PREC_FP32;
  const int nsplit = prec == B2P_PREC_BF16X3 ? 2 : 1;
  auto hi_ptr = [&](int id) -> __nv_bfloat16* {
    return id < 0 ? nullptr : reinterpret_cast<__nv_bfloat16*>(h->d_act + h->bufs[id].off * (size_t)h->cap);
  };
  auto lo_ptr = [&](int id) -> __nv_bfloat16* {
    return (id < 0 || nsplit != 2) ? nullptr : hi_ptr(id) + (size_t)h->bufs[id].L * h->bufs[id].C * h->cap;
  };
  bool proj_done = false;   // residual projection of the raw trajectory already produced by the first conv launch
  // ---- row-owned chain kernel for the 64-channel layers (tensor-core precisions) ----
  l2_weight_window() = L2Window{h->d_pack16, tc ? h->l2_window_bytes : 0};
  const bool chain = tc && h->chain_ok && h->chain_on;
  const int u0 = h->chain_u0;
  if (seam && !chain) return h->fail(B2P_ERR_STATE, "seam fusion needs the chain kernel");
  auto launch_chain = [&](bool do_tail, bool do_head) -> int {
    ChainArgs ca;
    memset(&ca, 0, sizeof(ca));
    auto mk = [&](int oi2, int kind, int T, int pad, int L, int in_buf, int out_buf, int res_kind, int res_buf, int phase, int ksteps) {
      const LayerOp& o = h->ops[oi2];
      ChainOp& c = ca.ops[ca.n_ops++];
      c.kind = kind; c.T = T; c.pad = pad; c.L = L; c.log2L = ilog2(L); c.in_buf = in_buf; c.out_buf = out_buf; c.res_kind = res_kind; c.res_buf = res_buf;
      c.temb_off = o.temb_off; c.phase = phase; c.gn = o.gamma != NPOS; c.ksteps = ksteps; c.w_off = h->chain_woff[oi2];
      c.bias = P + o.bias; c.gamma = o.gamma != NPOS ? P + o.gamma : nullptr; c.beta = o.gamma != NPOS ? P + o.beta : nullptr;
    };
    const int nops = (int)h->ops.size();
    if (do_tail) {
      mk(u0 + 1, CH_CONV, 5, 2, 8, 0, 1, CH_RES_F32, 0, 0, 4);
      mk(u0 + 2, CH_CONV, 5, 2, 8, 1, 2, CH_RES_NONE, 0, 0, 4);
      mk(u0 + 3, CH_CONV, 5, 2, 8, 2, 0, CH_RES_SMEM, 1, 0, 4);
      mk(u0 + 4, CH_UP, 4, 1, 8, 0, 1, CH_RES_NONE, 0, 0, 4);
      mk(u0 + 5, CH_CONV, 5, 2, 16, 1, CH_OUT_HEAD, CH_RES_NONE, 0, 0, 4);
      const LayerOp& hd = h->ops[nops - 1];
      ca.in_hi = hi_ptr(h->ops[u0].out); ca.in_lo = lo_ptr(h->ops[u0].out); ca.res_f32 = h->d_res0;
      ca.headW = P + hd.headW; ca.headB = P + hd.headB; ca.head_dim = hd.head_dim; ca.head_out = seam ? nullptr : head_out;
      if (seam) {
        ca.do_sched = 1;
        int rc2 = sched_make_args(seam->sched, &ca.sk);
        if (rc2) return rc2;
        ca.x_out = seam->sched.prev;
      }
    }
    if (do_head) {
      const int ph = do_tail ? 1 : 0;
      mk(0, CH_CONV, 1, 0, 16, CH_IN_IM2COL, 1, CH_RES_NONE, 0, ph, (5 * h->D + 15) / 16);
      mk(1, CH_CONV, 5, 2, 16, 1, 2, CH_RES_XPROJ, 0, ph, 4);
      mk(2, CH_CONV, 5, 2, 16, 2, 1, CH_RES_NONE, 0, ph, 4);
      mk(3, CH_CONV, 5, 2, 16, 1, 0, CH_RES_SMEM, 2, ph, 4);
      mk(4, CH_DOWN, 3, 1, 16, 0, CH_OUT_GLOBAL, CH_RES_NONE, 0, ph, 4);
      ca.xprojW = P + h->ops[1].resW; ca.xprojB = P + h->ops[1].resB;
      ca.out_hi = hi_ptr(h->ops[4].out); ca.out_lo = lo_ptr(h->ops[4].out);
    }
    ca.B = rows; ca.H = h->H; ca.D = h->D; ca.wpack = h->d_chain;
    ca.ns = 8;   // full 128-row tiles.  Half tiles (4 trajectories, twice the CTAs) measured SLOWER: a warp's TMEM lane quadrant is its SM sub-partition, so rows 64..127 idle two of the four schedulers
    ca.x = x; ca.x_period = x_period;
    ca.temb_rows = temb_rows; ca.temb_stride = h->temb_total;
    ca.temb2[0] = ttab_row; ca.temb2[1] = seam ? seam->next_ttab_row : nullptr;
    ca.trace = (h->d_chain_trace && do_tail && do_head) ? h->d_chain_trace : nullptr;
    int rc2 = launch_chain64(ca, nsplit, s);
    if (rc2) return h->fail(rc2, "chain kernel launch failed");
    ++*launches;
    return B2P_OK;
  };
  for (size_t oi = 0; oi < h->ops.size(); ++oi) {
    const LayerOp& op = h->ops[oi];
    if (chain && oi == 0) {            // ops 0..4: the head chain (already run by the previous step's seam launch inside a plan)
      if (!head_chain_done && (rc = launch_chain(false, true))) return rc;
      oi = 4;
      continue;
    }
    if (chain && (int)oi == u0 + 1) {  // ops u0+1 .. last: the tail chain (+ scheduler step + the next evaluation's head chain)
      if ((rc = launch_chain(true, seam && seam->next_head))) return rc;
      break;
    }
    if (tc && op.tcW_hi != NPOS) {
      // ------------------------------ tcgen05 path ------------------------------
      TcArgs t;
      TcMaps m;
      memset(&t, 0, sizeof(t));
      memset(&m, 0, sizeof(m));
      t.C[0] = op.C0; t.C[1] = op.C1; t.Cout = op.Cout;
      const int Lrows = op.Lin;   // GEMM rows are INPUT positions; taps are combined as row shifts in the epilogue
      t.n_out = 1; t.out_ldiv = 1; t.out_lmul = 1; t.out_L = op.Lout;
      if (!op.transposed) {
        int jmin = 0, jmax = op.taps - 1;
        if (op.stride == 1) {   // taps that can reach a valid input position (k=5 on L=2 only touches taps 1..3)
          jmin = op.pad - (op.Lout - 1) > 0 ? op.pad - (op.Lout - 1) : 0;
          jmax = op.pad + op.Lin - 1 < op.taps - 1 ? op.pad + op.Lin - 1 : op.taps - 1;
        }
        t.T = jmax - jmin + 1; t.tap0 = jmin; t.nt[0] = t.T;
        for (int i = 0; i < t.T; ++i) { t.tap_blk[0][i] = i; t.tap_shift[0][i] = jmin + i - op.pad; }
        t.out_ldiv = op.stride;   // stride-2 conv: out[lo] = sum_j Y_j[2*lo + j - pad], emitted by the even rows
      } else {
        // ConvTranspose1d(k4, s2, p1): out[2m] = Y1[m] + Y3[m-1];  out[2m+1] = Y0[m+1] + Y2[m]
        t.T = 4; t.tap0 = 0; t.n_out = 2; t.out_lmul = 2;
        t.nt[0] = 2; t.tap_blk[0][0] = 1; t.tap_shift[0][0] = 0; t.tap_blk[0][1] = 3; t.tap_shift[0][1] = -1;
        t.nt[1] = 2; t.tap_blk[1][0] = 0; t.tap_shift[1][0] = 1; t.tap_blk[1][1] = 2; t.tap_shift[1][1] = 0;
      }
      const bool aux = chain && (int)oi == u0;   // the block's residual projection rides along as a sixth tap (consumed by the tail chain)
      if (aux) { t.T = 6; t.aux_blk = 5; t.aux_out = h->d_res0; t.aux_bias = P + h->ops[u0 + 1].resB; }
      const int lstride = 1;
      t.Lrows = Lrows; t.log2L = ilog2(Lrows); t.samples_per_tile = 128 / Lrows; t.nrows = rows * Lrows;
      if (op.gamma != NPOS) { t.gn_gamma = P + op.gamma; t.gn_beta = P + op.beta; t.cg = op.Cout / 8; }
      if (op.headW != NPOS) { t.headW = P + op.headW; t.headB = P + op.headB; t.head_dim = op.head_dim; t.head_out = head_out; }
      if ((rc = tc_configure(t))) return h->fail(rc, "unsupported layer tiling");
      // with activation multicast (cluster_l > cluster_n) each CTA of the cluster loads 1/cluster_l of the row tile's samples
      const int box_b = t.cluster_l > t.cluster_n ? t.samples_per_tile / t.cluster_l : t.samples_per_tile;
      const int ins[2] = {op.in0, op.in1};
      const int cs[2] = {op.C0, op.C1};
      for (int sidx = 0; sidx < 2; ++sidx) {
        if (cs[sidx] == 0) continue;
        if ((rc = tc_make_act_map(&m.a[sidx][0], hi_ptr(ins[sidx]), rows, op.Lin, cs[sidx], op.Lin, lstride, box_b))) return h->fail(rc, "tensor map (A)");
        if (nsplit == 2 && (rc = tc_make_act_map(&m.a[sidx][1], lo_ptr(ins[sidx]), rows, op.Lin, cs[sidx], op.Lin, lstride, box_b))) return h->fail(rc, "tensor map (A lo)");
      }
      const __nv_bfloat16* P16 = reinterpret_cast<const __nv_bfloat16*>(h->d_pack16);
      const int wbox_t = t.cluster_m > 1 ? 1 : t.T;   // weight multicast: one box per tap, dealt to the CTAs of the cluster
      if ((rc = tc_make_weight_map(&m.w[0], P16 + (aux ? h->aux_hi : op.tcW_hi), aux ? 6 : op.taps, op.Cout, op.C0 + op.C1, wbox_t, t.tile_n))) return h->fail(rc, "tensor map (W)");
      if (nsplit == 2 && (rc = tc_make_weight_map(&m.w[1], P16 + (aux ? h->aux_lo : op.tcW_lo), aux ? 6 : op.taps, op.Cout, op.C0 + op.C1, wbox_t, t.tile_n))) return h->fail(rc, "tensor map (W lo)");
      if (op.resW != NPOS) {
        if (op.tcRW_hi != NPOS) {
          t.RC[0] = op.RC0; t.RC[1] = op.RC1; t.resB = P + op.resB;
          const int rins[2] = {op.rin0, op.rin1};
          const int rcs[2] = {op.RC0, op.RC1};
          for (int sidx = 0; sidx < 2; ++sidx) {
            if (rcs[sidx] == 0) continue;
            if ((rc = tc_make_act_map(&m.r[sidx][0], hi_ptr(rins[sidx]), rows, op.Lout, rcs[sidx], op.Lout, 1, box_b))) return h->fail(rc, "tensor map (R)");
            if (nsplit == 2 && (rc = tc_make_act_map(&m.r[sidx][1], lo_ptr(rins[sidx]), rows, op.Lout, rcs[sidx], op.Lout, 1, box_b))) return h->fail(rc, "tensor map (R lo)");
          }
          if ((rc = tc_make_weight_map(&m.rw[0], P16 + op.tcRW_hi, 1, op.Cout, op.RC0 + op.RC1, 0, t.tile_n))) return h->fail(rc, "tensor map (RW)");
          if (nsplit == 2 && (rc = tc_make_weight_map(&m.rw[1], P16 + op.tcRW_lo, 1, op.Cout, op.RC0 + op.RC1, 0, t.tile_n))) return h->fail(rc, "tensor map (RW lo)");
        } else if (proj_done) {
          t.res_f32 = h->d_res0;
        } else {
          // residual projection of the raw trajectory (C_in = transition_dim): tiny fp32 1x1 conv on CUDA cores
          ConvArgs pr{};
          pr.x0 = buf_ptr(h, op.rin0, x); pr.C0 = op.RC0; pr.x0_period = (op.rin0 == BUF_X) ? x_period : 0;
          pr.Lin = pr.Lout = op.Lout; pr.log2Lout = ilog2(op.Lout); pr.nrows = rows * op.Lout; pr.Cout = op.Cout;
          pr.taps = 1; pr.jmin = pr.jmax = 0; pr.stride = 1; pr.W = P + op.resW; pr.bias = P + op.resB; pr.out = h->d_res0;
          if ((rc = launch_conv_ffma(pr, s))) return h->fail(rc, "residual projection launch failed");
          ++*launches;
          t.res_f32 = h->d_res0;
        }
      } else if (op.res_id != BUF_NONE) {
        t.res_hi = hi_ptr(op.res_id); t.res_lo = lo_ptr(op.res_id);
      }
      t.bias = P + op.bias;
      if (op.gamma != NPOS) { t.gn_gamma = P + op.gamma; t.gn_beta = P + op.beta; t.cg = op.Cout / 8; }
      if (op.temb_off >= 0) { t.temb = temb_rows + op.temb_off; t.temb_stride = h->temb_total; t.temb2 = ttab_row ? ttab_row + op.temb_off : nullptr; }
      if (op.headW != NPOS) { t.headW = P + op.headW; t.headB = P + op.headB; t.head_dim = op.head_dim; t.head_out = head_out; }
      if (op.headW == NPOS) { t.out_hi = hi_ptr(op.out); t.out_lo = lo_ptr(op.out); }   // the head layer's own output is not materialised
      {
        static int tma_out = -1;
        if (tma_out < 0) { const char* e = getenv("B2P_TC_TMAOUT"); tma_out = e ? atoi(e) : 1; }
        if (tma_out && t.out_hi && t.n_out == 1 && t.out_ldiv == 1 && !t.headW && t.out_L == t.Lrows && t.out_lmul == 1) {
          if ((rc = tc_make_out_map(&m.o[0], t.out_hi, t.nrows, op.Cout, t.tile_n))) return h->fail(rc, "tensor map (out)");
          if (nsplit == 2 && (rc = tc_make_out_map(&m.o[1], t.out_lo, t.nrows, op.Cout, t.tile_n))) return h->fail(rc, "tensor map (out lo)");
          t.tma_out = 1;
        }
      }
      if ((rc = launch_conv_tc(m, t, nsplit, s))) return h->fail(rc, std::string("tcgen05 conv launch failed: ") + tc_last_error());
      ++*launches;
      continue;
    }
    ConvArgs a{};
    a.x0 = buf_ptr(h, op.in0, x); a.x1 = buf_ptr(h, op.in1, x); a.C0 = op.C0; a.C1 = op.C1;
    a.x0_period = (op.in0 == BUF_X) ? x_period : 0;
    a.Lin = op.Lin; a.Lout = op.Lout; a.log2Lout = ilog2(op.Lout); a.nrows = rows * op.Lout; a.Cout = op.Cout;
    a.taps = op.taps; a.stride = op.stride; a.pad = op.pad; a.transposed = op.transposed;
    // taps that can reach a valid input position for at least one output position
    a.jmin = 0; a.jmax = op.taps - 1;
    if (!op.transposed && op.stride == 1) {
      a.jmin = op.pad - (op.Lout - 1) > 0 ? op.pad - (op.Lout - 1) : 0;
      a.jmax = op.pad + op.Lin - 1 < op.taps - 1 ? op.pad + op.Lin - 1 : op.taps - 1;
    }
    a.W = P + op.W; a.bias = P + op.bias;
    if (op.gamma != NPOS) { a.gn_gamma = P + op.gamma; a.gn_beta = P + op.beta; a.cg = op.Cout / 8; }
    if (op.temb_off >= 0) { a.temb = temb_rows + op.temb_off; a.temb_stride = h->temb_total; a.temb2 = ttab_row ? ttab_row + op.temb_off : nullptr; }
    a.res_id = buf_ptr(h, op.res_id, x);
    if (op.resW != NPOS) {
      a.rx0 = buf_ptr(h, op.rin0, x); a.rx1 = buf_ptr(h, op.rin1, x); a.RC0 = op.RC0; a.RC1 = op.RC1;
      a.rx0_period = (op.rin0 == BUF_X) ? x_period : 0;
      a.resW = P + op.resW; a.resB = P + op.resB;
    }
    if (op.headW != NPOS) { a.headW = P + op.headW; a.headB = P + op.headB; a.head_dim = op.head_dim; a.head_out = head_out; }
    a.out = op.out == BUF_NONE ? nullptr : const_cast<float*>(buf_ptr(h, op.out, x));
    if (tc) {   // CUDA-core layer feeding tensor-core layers: emit bf16 hi/lo instead of fp32
      a.out = nullptr; a.out_hi = hi_ptr(op.out); a.out_lo = lo_ptr(op.out);
      if (oi + 1 < h->ops.size()) {   // also emit the next launch's 1x1 residual projection of the same input (fp32)
        const LayerOp& nx = h->ops[oi + 1];
        if (nx.resW != NPOS && nx.tcRW_hi == NPOS && nx.rin0 == op.in0 && nx.rin1 == BUF_NONE && nx.Lout == op.Lout && nx.Cout == op.Cout) {
          a.rx0 = buf_ptr(h, nx.rin0, x); a.RC0 = nx.RC0; a.rx0_period = (nx.rin0 == BUF_X) ? x_period : 0;
          a.resW = P + nx.resW; a.resB = P + nx.resB; a.res_out = h->d_res0;
          proj_done = true;
        }
      }
    }
    if (!tc && use_gemv) {
      // small batch: exact-fp32 GEMV kernels (one per layer; the fused 1x1 head becomes its own tiny launch)
      ConvArgs g = a;
      g.Wk = P + op.Wk; g.resWk = op.resWk != NPOS ? P + op.resWk : nullptr;
      g.headW = nullptr; g.headB = nullptr; g.head_out = nullptr;
      if (conv_gemv_applicable(g)) {
        if ((rc = launch_conv_gemv(g, s))) return h->fail(rc, "gemv conv launch failed");
        ++*launches;
        if (op.headW != NPOS) {
          ConvArgs hd{};
          hd.x0 = g.out; hd.C0 = op.Cout; hd.Lin = hd.Lout = op.Lout; hd.log2Lout = ilog2(op.Lout); hd.nrows = rows * op.Lout;
          hd.Cout = op.head_dim; hd.taps = 1; hd.jmin = hd.jmax = 0; hd.stride = 1; hd.Wk = P + op.headWk; hd.bias = P + op.headB;
          hd.out = head_out;
          if ((rc = launch_conv_gemv(hd, s))) return h->fail(rc, "gemv head launch failed");
          ++*launches;
        }
        continue;
      }
    }
    if (op.headW != NPOS) a.out = nullptr;   // fused head: the conv block's own output is not materialised
    if ((rc = launch_conv_ffma(a, s))) return h->fail(rc, "conv launch failed");
    ++*launches;
  }
  return B2P_OK;
}

// architecture tables and the state_dict slot table of a fresh handle (pure host code)
int init_handle_host(b2p_handle_s* h, const b2p_model_config* cfg) {
  h->cfg = *cfg;
  h->H = cfg->horizon; h->D = cfg->transition_dim; h->dim = cfg->dim; h->nlev = cfg->n_mults;
  h->chans[0] = h->D;
  h->temb_total = 0;
  for (int i = 0; i < h->nlev; ++i) {
    h->chans[i + 1] = cfg->dim * cfg->dim_mults[i];
    if (h->chans[i + 1] % 64 != 0) return B2P_ERR_INVALID_ARG;
  }
  for (int i = 0; i < h->nlev; ++i) h->temb_total += 2 * h->chans[i + 1];       // downs
  h->temb_total += 2 * h->chans[h->nlev];                                         // mid
  for (int u = 0; u < h->nlev - 1; ++u) h->temb_total += 2 * h->chans[h->nlev - 1 - u];  // ups
  build_slots(h);
  return B2P_OK;
}

}  // namespace

// ================================================= C ABI ====================================================
extern "C" {

int b2p_abi_version(void) { return B2P_ABI_VERSION; }

const char* b2p_status_string(int st) {
  switch (st) {
    case B2P_OK: return "ok";
    case B2P_ERR_INVALID_ARG: return "invalid argument";
    case B2P_ERR_UNKNOWN_WEIGHT: return "unknown state_dict key";
    case B2P_ERR_BAD_SHAPE: return "weight numel mismatch";
    case B2P_ERR_NOT_FINALIZED: return "weights not finalized";
    case B2P_ERR_MISSING_WEIGHT: return "missing weights";
    case B2P_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case B2P_ERR_STATE: return "invalid call sequence";
    default: return st > 0 ? cudaGetErrorString((cudaError_t)st) : "unknown";
  }
}

const char* b2p_last_error(b2p_handle h) { return h ? h->err.c_str() : ""; }

int b2p_create(const b2p_model_config* cfg, int device, b2p_handle* out) {
  if (!cfg || !out) return B2P_ERR_INVALID_ARG;
  if (cfg->n_mults < 2 || cfg->n_mults > 6 || cfg->dim % 64 != 0 || cfg->dim > 256 || cfg->transition_dim < 4 ||
      cfg->transition_dim > 16 || cfg->guidance < 0 || cfg->guidance > 2)
    return B2P_ERR_INVALID_ARG;
  int down = 1 << (cfg->n_mults - 1);
  if (cfg->horizon % down != 0 || cfg->horizon > 64 || (cfg->horizon & (cfg->horizon - 1)) != 0) return B2P_ERR_INVALID_ARG;
  if ((cfg->horizon * cfg->transition_dim) % 4 != 0) return B2P_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return B2P_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  B2P_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return B2P_ERR_NO_DEVICE;  // sm_100a cubin only: no fallback
  B2P_CUDA_TRY(cudaSetDevice(device));
  b2p_handle_s* h = new b2p_handle_s();
  h->device = device;
  if (int rc = init_handle_host(h, cfg)) { delete h; return rc; }
  {  // sinusoidal frequencies, fp32 as torch computes them (modeling/helpers.py:69-71)
    int half = h->dim / 2;
    std::vector<float> f(half);
    float sc = (float)(-(log(10000.0) / (half - 1)));
    for (int i = 0; i < half; ++i) f[i] = expf((float)i * sc);
    int rc = upload_freq_table(f.data(), half);
    if (rc) { delete h; return rc; }
  }
  cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
  h->chain_on = chain_enabled();
  if (getenv("B2P_CHAIN_TRACE")) { cudaMalloc((void**)&h->d_chain_trace, 8 * 16 * 16); cudaMemset(h->d_chain_trace, 0, 8 * 16 * 16); }
  *out = h;
  return B2P_OK;
}

int b2p_destroy(b2p_handle h) {
  if (!h) return B2P_OK;
  cudaSetDevice(h->device);
  drop_graphs(h);
  if (h->d_pack) cudaFree(h->d_pack);
  if (h->d_pack16) cudaFree(h->d_pack16);
  if (h->d_chain) cudaFree(h->d_chain);
  if (h->d_ws) cudaFree(h->d_ws);
  if (h->p_x) cudaFree(h->p_x);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
  return B2P_OK;
}

int b2p_num_weights(b2p_handle h) { return h ? (int)h->slots.size() : 0; }

int b2p_weight_info(b2p_handle h, int i, const char** key, int64_t* numel) {
  if (!h || i < 0 || i >= (int)h->slots.size()) return B2P_ERR_INVALID_ARG;
  if (key) *key = h->slots[i].key.c_str();
  if (numel) *numel = h->slots[i].numel;
  return B2P_OK;
}

int b2p_load_weight(b2p_handle h, const char* key, const float* data, int64_t numel) {
  if (!h || !key || !data) return B2P_ERR_INVALID_ARG;
  auto it = h->index.find(key);
  if (it == h->index.end()) return h->fail(B2P_ERR_UNKNOWN_WEIGHT, std::string("unknown key ") + key);
  Slot& s = h->slots[it->second];
  if (s.numel != numel) return h->fail(B2P_ERR_BAD_SHAPE, std::string("numel mismatch for ") + key);
  s.host.assign(data, data + numel);
  s.set = true;
  h->finalized = false;
  return B2P_OK;
}

int b2p_finalize_weights(b2p_handle h) {
  if (!h) return B2P_ERR_INVALID_ARG;
  for (auto& s : h->slots)
    if (!s.set) return h->fail(B2P_ERR_MISSING_WEIGHT, "missing key " + s.key);
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  int rc = build_program(h);
  if (rc) return h->fail(rc, "unsupported architecture");
  build_chain(h);
  std::vector<size_t> tp_offs;
  if (h->cfg.guidance == B2P_CLASSIFIER_GUIDANCE) {
    Packer pk{h->pack_host};
    build_trajpred(h, pk, tp_offs);
  }
  drop_graphs(h);
  if (h->d_pack) { B2P_CUDA_TRY(cudaFree(h->d_pack)); h->d_pack = nullptr; }
  B2P_CUDA_TRY(cudaMalloc((void**)&h->d_pack, h->pack_host.size() * sizeof(float)));
  B2P_CUDA_TRY(cudaMemcpy(h->d_pack, h->pack_host.data(), h->pack_host.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (h->d_chain) { B2P_CUDA_TRY(cudaFree(h->d_chain)); h->d_chain = nullptr; }
  if (h->chain_ok) {
    B2P_CUDA_TRY(cudaMalloc((void**)&h->d_chain, h->chain_host.size() + 1024));
    B2P_CUDA_TRY(cudaMemcpy(h->d_chain, h->chain_host.data(), h->chain_host.size(), cudaMemcpyHostToDevice));
  }
  if (h->d_pack16) { B2P_CUDA_TRY(cudaFree(h->d_pack16)); h->d_pack16 = nullptr; }
  B2P_CUDA_TRY(cudaMalloc((void**)&h->d_pack16, h->pack16_host.size() * sizeof(uint16_t) + 256));
  B2P_CUDA_TRY(cudaMemcpy(h->d_pack16, h->pack16_host.data(), h->pack16_host.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  {  // L2 residency of the tensor-core weight pack: reserve persisting lines and remember the window (applied per launch)
    static int want = -1;
    if (want < 0) { const char* e = getenv("B2P_L2_WINDOW"); want = e ? atoi(e) : 1; }
    h->l2_window_bytes = 0;
    if (want) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        size_t bytes = h->pack16_host.size() * sizeof(uint16_t);
        if (bytes > (size_t)prop.accessPolicyMaxWindowSize) bytes = (size_t)prop.accessPolicyMaxWindowSize;
        size_t lim = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
        if (getenv("B2P_L2_DEBUG")) fprintf(stderr, "[b2p] L2 %d B, persisting max %d B, window max %d B, pack16 %zu B\n", prop.l2CacheSize, prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize, bytes);
        if (getenv("B2P_L2_CARVE_MB")) lim = (size_t)atoi(getenv("B2P_L2_CARVE_MB")) << 20;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) h->l2_window_bytes = bytes;
        else cudaGetLastError();
      }
    }
  }
  if (!tp_offs.empty()) resolve_trajpred(h, tp_offs);
  // workspace offsets depend on the buffer table: force re-allocation
  if (h->d_ws) { B2P_CUDA_TRY(cudaFree(h->d_ws)); h->d_ws = nullptr; h->cap = 0; }
  h->finalized = true;
  return B2P_OK;
}

int b2p_set_precision(b2p_handle h, int precision) {
  if (!h || precision < 0 || precision > 2) return B2P_ERR_INVALID_ARG;
  h->cfg.precision = precision;
  drop_graphs(h);
  return B2P_OK;
}

// developer: stage clocks of CTA 0 of the last seam launch ([op][8] cycle counters; needs B2P_CHAIN_TRACE=1 at b2p_create)
__attribute__((visibility("default"))) int b2p_debug_chain_trace(b2p_handle h, unsigned long long* out) {
  if (!h || !out || !h->d_chain_trace) return B2P_ERR_STATE;
  return (int)cudaMemcpy(out, h->d_chain_trace, 8 * 16 * 16, cudaMemcpyDeviceToHost);
}

int b2p_set_chain(b2p_handle h, int enabled) {
  if (!h) return B2P_ERR_INVALID_ARG;
  h->chain_on = enabled != 0;
  drop_graphs(h);
  return B2P_OK;
}

int b2p_set_small_batch_max(b2p_handle h, int max_samples) {
  if (!h || max_samples < 0) return B2P_ERR_INVALID_ARG;
  h->small_batch_max = max_samples;
  drop_graphs(h);
  return B2P_OK;
}

int b2p_set_noise_seed(b2p_handle h, uint64_t seed) {
  if (!h) return B2P_ERR_INVALID_ARG;
  h->noise_seed = seed; h->noise_calls = 0;
  return B2P_OK;
}

uint64_t b2p_last_noise_key(b2p_handle h) { return h ? h->last_noise_key : 0; }

int64_t b2p_last_launch_count(b2p_handle h) { return h ? h->last_launches : 0; }
int64_t b2p_unet_flops_per_sample(b2p_handle h) { return h ? h->flops_per_sample : 0; }
int64_t b2p_weight_bytes(b2p_handle h) { return h ? (int64_t)(h->pack_host.size() * sizeof(float)) : 0; }

int b2p_unet_forward(b2p_handle h, const float* x, const float* feat, int32_t feat_rows, const int64_t* t, int32_t t_count,
                     const float* cond, float* out, float* action_out, float* time_embed_out, int32_t B, void* stream) {
  if (!h || !x || !feat || !t || B <= 0) return B2P_ERR_INVALID_ARG;
  if (feat_rows <= 0 || B % feat_rows != 0 || t_count <= 0 || B % t_count != 0) return h->fail(B2P_ERR_INVALID_ARG, "feat/t rows must divide B");
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  int64_t n = 0;
  int rc;
  if (h->cfg.guidance != B2P_CLASSIFIER_GUIDANCE) {
    if (!out) return B2P_ERR_INVALID_ARG;
    rc = run_unet(h, x, 0, feat, feat_rows, t, t_count, cond, out, time_embed_out, B, s, &n);
  } else {
    rc = ensure_workspace(h, B);
    if (rc) return rc;
    // action lands in a scratch [B,H,3] unless the caller wants it
    float* act = action_out;
    float* scratch = nullptr;
    if (!act) { B2P_CUDA_TRY(cudaMallocAsync((void**)&scratch, sizeof(float) * B * h->H * 3, s)); act = scratch; }
    float* te = time_embed_out ? time_embed_out : h->d_time_embed;
    rc = run_unet(h, x, 0, feat, feat_rows, t, t_count, cond, act, te, B, s, &n);
    if (!rc && out) {
      // out = cat[ cat[0, state_pred(action[:, :-1], te)], action ]  (modeling/temporal.py:237-241)
      rc = launch_state_pred(h->tp, act, te, h->dim, out, 1, B, h->H, h->D, s);
      ++n;
    }
    if (scratch) B2P_CUDA_TRY(cudaFreeAsync(scratch, s));
  }
  h->last_launches = n;
  return rc;
}

int b2p_state_pred(b2p_handle h, const float* action, const float* time_embed, float* state, int32_t B, void* stream) {
  if (!h || !action || !time_embed || !state || B <= 0) return B2P_ERR_INVALID_ARG;
  if (!h->finalized || !h->has_tp) return h->fail(B2P_ERR_STATE, "model has no state predictor");
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  // state [B, H-1, D-3]: dense rows, no zero row, no action columns
  return launch_state_pred(h->tp, action, time_embed, h->dim, state, 0, B, h->H, h->D, (cudaStream_t)stream);
}

int b2p_state_pred_vjp(b2p_handle h, const float* action, const float* time_embed, const float* grad_state, float* grad_action,
                       int32_t B, void* stream) {
  if (!h || !action || !time_embed || !grad_state || !grad_action || B <= 0) return B2P_ERR_INVALID_ARG;
  if (!h->finalized || !h->has_tp) return h->fail(B2P_ERR_STATE, "model has no state predictor");
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  return launch_state_pred_vjp(h->tp, action, time_embed, h->dim, grad_state, grad_action, B, h->H, h->D, (cudaStream_t)stream);
}

int b2p_classifier_guidance(b2p_handle h, float* model_output, const float* time_embed, const float* target, float grad_scale,
                            float classifier_scale, int32_t B, void* stream) {
  if (!h || !model_output || !time_embed || !target || B <= 0) return B2P_ERR_INVALID_ARG;
  if (!h->finalized || !h->has_tp) return h->fail(B2P_ERR_STATE, "model has no state predictor");
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  return launch_classifier_guidance(h->tp, model_output, time_embed, h->dim, target, grad_scale, classifier_scale, B, h->H, h->D,
                                    (cudaStream_t)stream);
}

// ------------------------------------------------ plan ---------------------------------------------------
static int ensure_plan_buffers(b2p_handle h, int B, int T) {
  if (B <= h->plan_capB && T <= h->plan_capT) return B2P_OK;
  drop_graphs(h);
  if (h->p_x) { B2P_CUDA_TRY(cudaFree(h->p_x)); h->p_x = nullptr; }
  int cb = B > h->plan_capB ? B : h->plan_capB, ct = T > h->plan_capT ? T : h->plan_capT;
  size_t hd = (size_t)h->H * h->D;
  size_t floats = (size_t)cb * (hd * 6 + h->dim + 2 + 4 + 1 + (size_t)h->H * 3 + h->temb_total) + (size_t)ct * (cb * hd + h->temb_total + h->dim) + 64 * 20;
  floats = (floats + 3) & ~(size_t)3;   // the int64 timestep table and the Philox key follow the floats
  B2P_CUDA_TRY(cudaMalloc((void**)&h->p_x, floats * sizeof(float) + sizeof(int64_t) * (ct + 8)));
  float* p = h->p_x;
  auto take = [&](size_t n) { float* r = p; p += (n + 63) & ~(size_t)63; return r; };
  take(cb * hd);  // p_x itself
  h->p_feat = take((size_t)cb * h->dim);
  h->p_target = take((size_t)cb * 2);
  h->p_cond = take((size_t)cb * 4);
  h->p_traj = take(cb * hd);
  h->p_mask = take(cb * hd);
  h->p_mo = take(2 * cb * hd);
  h->p_action = take((size_t)cb * h->H * 3);
  h->p_out = take(cb * hd);
  h->p_itab = take((size_t)cb * h->temb_total);
  h->p_ttab = take((size_t)ct * h->temb_total);
  h->p_tetab = take((size_t)ct * h->dim);
  h->p_thr = take((size_t)cb);
  h->p_noise = take((size_t)ct * cb * hd);
  h->p_tsteps = reinterpret_cast<int64_t*>(h->p_x + floats);
  h->p_seed = reinterpret_cast<unsigned long long*>(h->p_tsteps + ct + 1);
  h->ts_ntrain = h->ts_T = -1;
  h->plan_capB = cb; h->plan_capT = ct;
  return B2P_OK;
}

// enqueue the T-step loop on s, reading/writing the handle's static plan buffers
static int enqueue_plan(b2p_handle h, const b2p_plan_config& pc, const GraphKey& k, cudaStream_t s, int64_t* launches) {
  const int B = k.B, T = k.T;
  const size_t hd = (size_t)h->H * h->D;
  std::vector<float> ac(pc.sched.num_train_timesteps);
  static const char* kSched[3] = {"squaredcos_cap_v2", "linear", "scaled_linear"};
  if (pc.sched.beta_schedule < 0 || pc.sched.beta_schedule > 2) return B2P_ERR_INVALID_ARG;
  int rc = b2p_alphas_cumprod(kSched[pc.sched.beta_schedule], pc.sched.num_train_timesteps, pc.sched.beta_start,
                              pc.sched.beta_end, ac.data());
  if (rc) return rc;
  const int g = h->cfg.guidance;
  const bool inpaint = pc.sched.kind == B2P_SCHED_INPAINT_DDIM || pc.sched.kind == B2P_SCHED_INPAINT_DDPM;
  const bool tables = g != B2P_FREE_GUIDANCE;   // CFG adds cond_mlp(cond) to the time embedding per sample: no shared time vector
  if (tables) {
    // once per plan: time table [T, temb_total] (+ bias), image term [B, temb_total], time embeddings [T, dim]
    const float* P = h->d_pack;
    const int rows_e = T > B ? T : B;
    EmbedArgs e{};
    e.t = h->p_tsteps; e.t_count = T; e.feat = h->p_feat; e.feat_rows = B; e.use_cond = 0;
    e.w1t = P + h->o_w1t; e.b1 = P + h->o_b1; e.w3t = P + h->o_w3t; e.b3 = P + h->o_b3;
    e.time_embed = h->p_tetab; e.mish_te = h->d_mish_te; e.mish_feat = h->d_mish_feat; e.te_rows = T; e.feat_out_rows = B;
    e.B = rows_e; e.dim = h->dim;
    if ((rc = launch_embed(e, s))) return rc;
    ConvArgs a{};
    a.x0 = h->d_mish_te; a.C0 = h->dim; a.Lin = a.Lout = 1; a.nrows = T; a.Cout = h->temb_total; a.taps = 1; a.stride = 1;
    a.W = P + h->o_tembW; a.bias = P + h->o_tembB; a.out = h->p_ttab;
    if ((rc = launch_conv_ffma(a, s))) return rc;
    ConvArgs b{};
    b.x0 = h->d_mish_feat; b.C0 = h->dim; b.Lin = b.Lout = 1; b.nrows = B; b.Cout = h->temb_total; b.taps = 1; b.stride = 1;
    b.W = P + h->o_tembW + (size_t)h->dim * h->temb_total; b.out = h->p_itab;
    if ((rc = launch_conv_ffma(b, s))) return rc;
    *launches += 3;
  }
  // Without guidance the tail chain of evaluation i, the scheduler step and the head chain of evaluation i+1 are ONE launch
  // (chain64.cu): x_t stays in shared memory across the step boundary.
  const bool seam_plan = g == B2P_NO_GUIDANCE && h->chain_ok && h->chain_on && h->cfg.precision != B2P_PREC_FP32 && B > h->small_batch_max &&
                         !(pc.sched.thresholding && pc.sched.sample_max_value != 1.0f);
  for (int i = 0; i < T; ++i) {
    int t = (T - 1 - i) * (pc.sched.num_train_timesteps / T);
    b2p_step_coeffs kc;
    if ((rc = b2p_step_coeffs_compute(&pc.sched, ac.data(), T, t, pc.eta, &kc))) return rc;
    const int64_t* tp = h->p_tsteps + i;
    const float* mo = h->p_mo;
    const float* mo_u = nullptr;
    int flags = B2P_STEP_ZERO_FIRST_WAYPOINT;
    bool last = (i == T - 1);
    if (last && pc.postprocess) flags |= B2P_STEP_FINAL_POSTPROCESS;
    auto sched_launch = [&]() {
      return SchedLaunch{pc.sched, kc, mo, mo_u, pc.free_scale, h->p_x, k.has_noise ? h->p_noise + (size_t)i * B * hd : nullptr,
                         (inpaint && k.has_traj) ? h->p_traj : nullptr, (inpaint && k.has_mask) ? h->p_mask : nullptr,
                         last ? h->p_out : h->p_x, nullptr, B, h->H, h->D, pc.eta, pc.magic_num, flags,
                         k.dev_noise ? h->p_seed : nullptr, (unsigned)i, h->p_thr};
    };
    if (g == B2P_FREE_GUIDANCE) {
      // rows [0,B) conditional, [B,2B) unconditional (interact.py:119-127, 134-141)
      if ((rc = run_unet(h, h->p_x, B, h->p_feat, B, tp, 1, h->p_cond, h->p_mo, nullptr, 2 * B, s, launches))) return rc;
      mo_u = h->p_mo + (size_t)B * hd;
    } else if (g == B2P_CLASSIFIER_GUIDANCE) {
      const float* te_row = h->p_tetab + (size_t)i * h->dim;   // one time embedding shared by the batch (stride 0)
      if ((rc = run_unet(h, h->p_x, 0, h->p_feat, B, tp, 1, nullptr, h->p_action, nullptr, B, s, launches, h->p_itab,
                         h->p_ttab + (size_t)i * h->temb_total))) return rc;
      if (!inpaint && k.has_target) {   // state predictor forward + guidance gradient + update in one launch
        if ((rc = launch_classifier_guidance_from_action(h->tp, h->p_action, h->p_mo, te_row, 0, h->p_target, kc.guidance_grad_scale,
                                                         pc.classifier_scale, B, h->H, h->D, s))) return rc;
      } else {
        if ((rc = launch_state_pred(h->tp, h->p_action, te_row, 0, h->p_mo, 1, B, h->H, h->D, s))) return rc;
      }
      ++*launches;
    } else if (seam_plan) {
      ChainSeam seam{sched_launch(), !last, last ? nullptr : h->p_ttab + (size_t)(i + 1) * h->temb_total};
      if ((rc = run_unet(h, h->p_x, 0, h->p_feat, B, tp, 1, nullptr, h->p_mo, nullptr, B, s, launches, h->p_itab,
                         h->p_ttab + (size_t)i * h->temb_total, &seam, i > 0))) return rc;
      continue;                         // the scheduler step ran inside the seam launch
    } else {
      if ((rc = run_unet(h, h->p_x, 0, h->p_feat, B, tp, 1, nullptr, h->p_mo, nullptr, B, s, launches, h->p_itab,
                         h->p_ttab + (size_t)i * h->temb_total))) return rc;
    }
    SchedLaunch L = sched_launch();
    if ((rc = launch_sched_step(L, s))) return rc;
    ++*launches;
  }
  return B2P_OK;
}

__global__ void zero_first_waypoint_kernel(float* x, int B, int HD) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 3) x[(size_t)(i / 3) * HD + (i % 3)] = 0.f;
}

// The denoiser kernels of a plan (captured or eager) read their timesteps from h->p_tsteps when they RUN, so the table must
// hold this plan's schedule before every launch — a cached graph of another (num_train_timesteps, T) may have run in between.
static int ensure_timesteps(b2p_handle h, int ntrain, int T, cudaStream_t s) {
  if (h->ts_ntrain == ntrain && h->ts_T == T) return B2P_OK;
  std::vector<int64_t> ts(T);
  b2p_timesteps(ntrain, T, ts.data());
  B2P_CUDA_TRY(cudaMemcpyAsync(h->p_tsteps, ts.data(), sizeof(int64_t) * T, cudaMemcpyHostToDevice, s));
  B2P_CUDA_TRY(cudaStreamSynchronize(s));   // ts is a stack vector
  h->ts_ntrain = ntrain; h->ts_T = T;
  return B2P_OK;
}

// host: the tensor arguments are HOST pointers (H2D / D2H copies on s).  sync: synchronise s before returning (a host caller
// without a stream of its own).  noise_batch: batch extent of the caller's noise tensor [T, noise_batch, H, D] when this call
// plans a contiguous slice [b0, b0 + B) of a larger batch (the caller passes noise + b0*H*D); 0 or B = dense.
static int plan_impl(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
                     const float* noise, const float* traj, const float* mask, float* out, int B, cudaStream_t s, bool host,
                     bool sync, int noise_batch = 0) {
  if (!h || !pc || !x_init || !feat || !out || B <= 0) return B2P_ERR_INVALID_ARG;
  if (!h->finalized) return h->fail(B2P_ERR_NOT_FINALIZED, "weights not finalized");
  const int T = pc->num_inference_steps;
  if (T <= 0 || T > pc->sched.num_train_timesteps) return h->fail(B2P_ERR_INVALID_ARG, "bad num_inference_steps");
  const int g = h->cfg.guidance;
  const bool ddpm = pc->sched.kind == B2P_SCHED_GUIDANCE_DDPM || pc->sched.kind == B2P_SCHED_INPAINT_DDPM;
  const bool inpaint = pc->sched.kind == B2P_SCHED_INPAINT_DDIM || pc->sched.kind == B2P_SCHED_INPAINT_DDPM;
  // a scheduler that consumes noise and gets none draws it in the kernel (Philox keyed by b2p_set_noise_seed + a per-plan counter)
  const bool dev_noise = (ddpm || (inpaint && traj && mask) || pc->eta > 0.f) && !noise;
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure_plan_buffers(h, B, T))) return rc;
  if ((rc = ensure_workspace(h, g == B2P_FREE_GUIDANCE ? 2 * B : (B > T ? B : T)))) return rc;
  const size_t hd = (size_t)h->H * h->D;
  const cudaMemcpyKind kind = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  B2P_CUDA_TRY(cudaMemcpyAsync(h->p_x, x_init, sizeof(float) * B * hd, kind, s));
  B2P_CUDA_TRY(cudaMemcpyAsync(h->p_feat, feat, sizeof(float) * B * h->dim, kind, s));
  if (target) {
    B2P_CUDA_TRY(cudaMemcpyAsync(h->p_target, target, sizeof(float) * B * 2, kind, s));
    if (g == B2P_FREE_GUIDANCE) {  // cond = cat[target, zeros]
      B2P_CUDA_TRY(cudaMemcpyAsync(h->p_cond, target, sizeof(float) * B * 2, kind, s));
      B2P_CUDA_TRY(cudaMemsetAsync(h->p_cond + (size_t)B * 2, 0, sizeof(float) * B * 2, s));
    }
  } else if (g == B2P_FREE_GUIDANCE) {   // generate_traj(image, None): cond = None == zeros for both halves (modeling/temporal.py:207)
    B2P_CUDA_TRY(cudaMemsetAsync(h->p_cond, 0, sizeof(float) * B * 4, s));
  }
  if (noise) {
    if (noise_batch > B)   // a batch slice of a larger [T, noise_batch, H, D] tensor: T strided rows
      B2P_CUDA_TRY(cudaMemcpy2DAsync(h->p_noise, sizeof(float) * B * hd, noise, sizeof(float) * noise_batch * hd, sizeof(float) * B * hd, T, kind, s));
    else
      B2P_CUDA_TRY(cudaMemcpyAsync(h->p_noise, noise, sizeof(float) * T * B * hd, kind, s));
  }
  if (traj) B2P_CUDA_TRY(cudaMemcpyAsync(h->p_traj, traj, sizeof(float) * B * hd, kind, s));
  if (mask) B2P_CUDA_TRY(cudaMemcpyAsync(h->p_mask, mask, sizeof(float) * B * hd, kind, s));
  zero_first_waypoint_kernel<<<(B * 3 + 127) / 128, 128, 0, s>>>(h->p_x, B, (int)hd);  // interact.py:129

  GraphKey key;
  memset(&key, 0, sizeof(key));
  key.B = B; key.T = T; key.kind = pc->sched.kind; key.has_target = target != nullptr; key.has_noise = noise != nullptr;
  key.has_traj = traj != nullptr; key.has_mask = mask != nullptr; key.dev_noise = dev_noise; key.pc = *pc;
  if (dev_noise) {   // splitmix64 of (seed, plan counter): every plan of the handle gets its own stream of draws
    unsigned long long z = h->noise_seed + 0x9E3779B97F4A7C15ULL * (++h->noise_calls);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    h->last_noise_key = z;
    B2P_CUDA_TRY(cudaMemcpyAsync(h->p_seed, &z, sizeof(z), cudaMemcpyHostToDevice, s));   // pageable source: staged before the call returns
  }
  int64_t launches = 1;
  if ((rc = ensure_timesteps(h, pc->sched.num_train_timesteps, T, s))) return rc;
  if (pc->use_graph) {
    GraphEntry* ge = nullptr;
    for (auto& e : h->graphs) if (e.key == key) { ge = &e; break; }
    if (!ge) {
      cudaGraph_t graph;
      int64_t nl = 0;
      B2P_CUDA_TRY(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
      rc = enqueue_plan(h, *pc, key, h->cap_stream, &nl);
      cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
      if (rc) { if (ce == cudaSuccess) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return (int)ce;
      cudaGraphExec_t exec;
      ce = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) return (int)ce;
      if (h->graphs.size() >= 16) {            // a replayable plan holds thousands of kernel nodes: keep the 16 most recent shapes
        cudaGraphExecDestroy(h->graphs.front().exec);
        h->graphs.erase(h->graphs.begin());
      }
      h->graphs.push_back(GraphEntry{key, exec, nl});
      ge = &h->graphs.back();
    } else if (ge != &h->graphs.back()) {      // most recently used last
      GraphEntry hit = *ge;
      h->graphs.erase(h->graphs.begin() + (ge - h->graphs.data()));
      h->graphs.push_back(hit);
      ge = &h->graphs.back();
    }
    B2P_CUDA_TRY(cudaGraphLaunch(ge->exec, s));
    launches += ge->launches;
  } else {
    if ((rc = enqueue_plan(h, *pc, key, s, &launches))) return rc;
  }
  B2P_CUDA_TRY(cudaMemcpyAsync(out, h->p_out, sizeof(float) * B * hd, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  if (sync) B2P_CUDA_TRY(cudaStreamSynchronize(s));
  h->last_launches = launches;
  return B2P_OK;
}

int b2p_plan(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
             const float* noise, const float* target_traj, const float* target_mask, float* out, int32_t B, void* stream) {
  return plan_impl(h, pc, x_init, feat, target, noise, target_traj, target_mask, out, B, (cudaStream_t)stream, false, false);
}

int b2p_plan_host(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
                  const float* noise, const float* target_traj, const float* target_mask, float* out, int32_t B) {
  if (!h) return B2P_ERR_INVALID_ARG;
  return plan_impl(h, pc, x_init, feat, target, noise, target_traj, target_mask, out, B, h->cap_stream, true, true);
}

int b2p_plan_host_async(b2p_handle h, const b2p_plan_config* pc, const float* x_init, const float* feat, const float* target,
                        const float* noise, int32_t noise_batch, const float* target_traj, const float* target_mask, float* out,
                        int32_t B) {
  if (!h) return B2P_ERR_INVALID_ARG;
  return plan_impl(h, pc, x_init, feat, target, noise, target_traj, target_mask, out, B, h->cap_stream, true, false, noise_batch);
}

int b2p_sync(b2p_handle h) {
  if (!h) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaSetDevice(h->device));
  B2P_CUDA_TRY(cudaStreamSynchronize(h->cap_stream));
  return B2P_OK;
}

// Independent planning requests sharded by batch over the GPUs of one box (the reference has a single caller
// with a single batch, interact.py:115-168): contiguous split, the first B % n shards take one extra trajectory; every
// handle's private stream gets its H2D copies, its captured loop and its D2H copy enqueued back to back, then all are
// joined.  No collective, no peer access: the only inter-device traffic is host staging.
int b2p_plan_sharded_host(const b2p_handle* handles, int32_t n_handles, const b2p_plan_config* pc, const float* x_init,
                          const float* feat, const float* target, const float* noise, const float* target_traj,
                          const float* target_mask, float* out, int32_t B) {
  if (!handles || n_handles <= 0 || !pc || !x_init || !feat || !out || B <= 0) return B2P_ERR_INVALID_ARG;
  for (int i = 0; i < n_handles; ++i) if (!handles[i]) return B2P_ERR_INVALID_ARG;
  const int q = B / n_handles, r = B % n_handles;
  int rc = B2P_OK, lo = 0;
  for (int i = 0; i < n_handles && !rc; ++i) {
    const int nb = q + (i < r ? 1 : 0);
    if (nb == 0) continue;
    b2p_handle h = handles[i];
    const size_t hd = (size_t)h->H * h->D;
    rc = plan_impl(h, pc, x_init + lo * hd, feat + (size_t)lo * h->dim, target ? target + (size_t)lo * 2 : nullptr,
                   noise ? noise + lo * hd : nullptr, target_traj ? target_traj + lo * hd : nullptr,
                   target_mask ? target_mask + lo * hd : nullptr, out + lo * hd, nb, h->cap_stream, true, false, B);
    lo += nb;
  }
  for (int i = 0; i < n_handles; ++i) {   // join every device even after an error so no copy is left in flight on caller memory
    cudaSetDevice(handles[i]->device);
    cudaError_t e = cudaStreamSynchronize(handles[i]->cap_stream);
    if (!rc && e != cudaSuccess) rc = (int)e;
  }
  return rc;
}

}  // extern "C"
