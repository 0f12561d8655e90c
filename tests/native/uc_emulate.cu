// CPU emulation of the EXPERIMENTAL one-cluster-per-trajectory denoiser kernel (csrc/unet_cluster.cu): runs the REAL host-side
// program builder (slot liveness, tap ranges, chunking, per-CTA weight streams: build_cluster_program_host in api.cu) and then
// walks the program exactly as the kernel does — per CTA rank, per chunk of its own stream, raw exchange, whole-tensor
// GroupNorm/Mish epilogue, head — in plain loops.  It checks everything about the kernel except its thread mapping and
// barriers, without a GPU.  Driven by tests/test_cluster_program.py, which supplies weights, one trajectory, the
// conditioning vector and the CPU oracle's output in a flat binary file.
//
// file: int32 n_keys, then per key {int32 len, bytes, int64 numel, float[numel]}, then x[H*D], cond_input[2*dim], expect[H*D]
#include "../../autonomous_driving_with_diffusion_model_b200/csrc/api.cu"

#include <cmath>
#include <cstdio>

static float mish_h(float x) {
  if (x > 20.f) return x;
  float e = expf(x);
  float n = e * (e + 2.f);
  return x * (n / (n + 2.f));
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: uc_emulate file\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  b2p_model_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.horizon = 16; cfg.transition_dim = 7; cfg.dim = 64; cfg.n_mults = 4;
  const int mults[4] = {1, 2, 4, 8};
  for (int i = 0; i < 4; ++i) cfg.dim_mults[i] = mults[i];
  b2p_handle_s* h = new b2p_handle_s();
  if (init_handle_host(h, &cfg)) { fprintf(stderr, "init failed\n"); return 1; }
  int32_t nk = 0;
  if (fread(&nk, 4, 1, f) != 1) return 2;
  for (int i = 0; i < nk; ++i) {
    int32_t len; int64_t numel;
    if (fread(&len, 4, 1, f) != 1) return 2;
    std::string key(len, ' ');
    if (fread(&key[0], 1, len, f) != (size_t)len || fread(&numel, 8, 1, f) != 1) return 2;
    std::vector<float> v(numel);
    if (fread(v.data(), 4, numel, f) != (size_t)numel) return 2;
    int rc = b2p_load_weight(h, key.c_str(), v.data(), numel);
    if (rc) { fprintf(stderr, "load %s: %d\n", key.c_str(), rc); return 1; }
  }
  const int HD = h->H * h->D;
  std::vector<float> x(HD), ci(2 * h->dim), expect(HD);
  if (fread(x.data(), 4, HD, f) != (size_t)HD || fread(ci.data(), 4, ci.size(), f) != ci.size() || fread(expect.data(), 4, HD, f) != (size_t)HD) return 2;
  fclose(f);
  for (auto& s : h->slots) if (!s.set) { fprintf(stderr, "missing %s\n", s.key.c_str()); return 1; }
  if (build_program(h)) { fprintf(stderr, "build_program failed\n"); return 1; }
  std::vector<UcProgram> pgv(1);
  UcProgram& pg = pgv[0];
  std::vector<float> stream;
  int rc = build_cluster_program_host(h, pg, stream);
  if (rc) { fprintf(stderr, "build_cluster_program_host: %d\n", rc); return 1; }
  const float* P = h->pack_host.data();
  const size_t S = pg.stream_floats_per_cta;

  // per-block time terms: Linear(Mish(cond_input)) for all 16 blocks at once (the [2*dim][temb_total] matrix of the pack)
  std::vector<float> temb(h->temb_total);
  for (int c = 0; c < h->temb_total; ++c) {
    float v = P[h->o_tembB + c];
    for (int i = 0; i < 2 * h->dim; ++i) v += mish_h(ci[i]) * P[h->o_tembW + (size_t)i * h->temb_total + c];
    temb[c] = v;
  }

  std::vector<float> slots((size_t)UC_NSLOT * UC_SLOT_FLOATS, 0.f), out(HD, 0.f);
  for (int i = 0; i < HD; ++i) slots[(size_t)pg.x_slot * UC_SLOT_FLOATS + i] = x[i];
  int q = 0, max_live_chunk_floats = 0;
  for (int oi = 0; oi < pg.n_ops; ++oi) {
    const UcOp& o = pg.ops[oi];
    std::vector<float> raw(UC_SLOT_FLOATS, 0.f), rraw(UC_SLOT_FLOATS, 0.f);
    if (o.chunk0 != q) { fprintf(stderr, "op %d: chunk cursor %d != chunk0 %d\n", oi, q, o.chunk0); return 1; }
    auto row_of = [&](int l, int jj) {   // input position feeding (output row l, tap jmin + jj), or -1
      const int j = o.jmin + jj;
      int pos;
      if (!o.transposed) pos = l * o.stride + j - o.pad;
      else { const int num = l + o.pad - j; pos = (num >= 0 && num % o.stride == 0) ? num / o.stride : -1; }
      return (pos >= 0 && pos < o.Lin) ? pos : -1;
    };
    for (int rank = 0; rank < UC_CL; ++rank)
      for (int cl = 0; cl < o.nc; ++cl) {
        const int ch = rank * o.nc + cl;
        for (int r = 0; r < o.Lout; ++r) {
          const int Cin = o.C0 + o.C1;
          float acc = 0.f;
          for (int c = 0; c < o.nchunks; ++c) {
            const UcChunk& ck = pg.chunks[o.chunk0 + c];
            if (ck.bytes != o.nc * ck.kstride * 4 || ck.bytes % 16 || ck.off % 4 || o.nc * ck.kstride > UC_STAGE_FLOATS) { fprintf(stderr, "bad chunk\n"); return 1; }
            if (o.nc * ck.kstride > max_live_chunk_floats) max_live_chunk_floats = o.nc * ck.kstride;
            for (int kk = 0; kk < ck.klen; ++kk) {
              const int k = ck.k0 + kk, jj = k / Cin, c_ = k % Cin;
              const int pos = row_of(r, jj);
              if (pos < 0) continue;
              const float xv = c_ < o.C0 ? slots[(size_t)o.in0 * UC_SLOT_FLOATS + pos * o.C0 + c_]
                                         : slots[(size_t)o.in1 * UC_SLOT_FLOATS + pos * o.C1 + (c_ - o.C0)];
              acc += stream[rank * S + ck.off + (size_t)cl * ck.kstride + kk] * xv;
            }
          }
          raw[r * o.Cout + ch] = acc + (o.bias >= 0 ? P[o.bias + ch] : 0.f);
          if (o.rnchunks > 0) {
            const int RCin = o.RC0 + o.RC1;
            float ra = 0.f;
            for (int c = 0; c < o.rnchunks; ++c) {
              const UcChunk& ck = pg.chunks[o.chunk0 + o.nchunks + c];
              for (int kk = 0; kk < ck.klen; ++kk) {
                const int k = ck.k0 + kk;
                if (k >= RCin) { fprintf(stderr, "bad residual chunk\n"); return 1; }
                const float xv = k < o.RC0 ? slots[(size_t)o.rin0 * UC_SLOT_FLOATS + r * o.RC0 + k]
                                           : slots[(size_t)o.rin1 * UC_SLOT_FLOATS + r * o.RC1 + (k - o.RC0)];
                ra += stream[rank * S + ck.off + (size_t)cl * ck.kstride + kk] * xv;
              }
            }
            rraw[r * o.Cout + ch] = ra + (o.resB >= 0 ? P[o.resB + ch] : 0.f);
          }
        }
      }
    q += o.nchunks + o.rnchunks;
    float mean[8] = {0}, rstd[8] = {0};
    if (o.gn) {
      const int cg = o.Cout / 8, n = o.Lout * cg;
      for (int g = 0; g < 8; ++g) {
        float s = 0.f;
        for (int r = 0; r < o.Lout; ++r) for (int c = 0; c < cg; ++c) s += raw[r * o.Cout + g * cg + c];
        const float m = s / n;
        float qq = 0.f;
        for (int r = 0; r < o.Lout; ++r) for (int c = 0; c < cg; ++c) { const float d = raw[r * o.Cout + g * cg + c] - m; qq += d * d; }
        mean[g] = m; rstd[g] = 1.0f / sqrtf(qq / n + 1e-5f);
      }
    }
    std::vector<float> res(UC_SLOT_FLOATS, 0.f);
    if (o.rnchunks > 0) res = rraw;
    else if (o.res_id >= 0) for (int e = 0; e < o.Lout * o.Cout; ++e) res[e] = slots[(size_t)o.res_id * UC_SLOT_FLOATS + e];
    if (o.out == o.in0 || o.out == o.in1 || o.out == o.res_id || o.out == o.rin0 || o.out == o.rin1) { fprintf(stderr, "op %d: output slot aliases an input\n", oi); return 1; }
    for (int e = 0; e < o.Lout * o.Cout; ++e) {
      const int ch = e % o.Cout;
      float v = raw[e];
      if (o.gn) { const int g = ch / (o.Cout / 8); v = mish_h((v - mean[g]) * rstd[g] * P[o.gamma + ch] + P[o.beta + ch]); }
      const float add = o.temb_off >= 0 ? temb[o.temb_off + ch] : 0.f;
      slots[(size_t)o.out * UC_SLOT_FLOATS + e] = v + add + res[e];
    }
    if (o.head)
      for (int r = 0; r < o.Lout; ++r)
        for (int j = 0; j < pg.head_dim; ++j) {
          float v = 0.f;
          for (int c = 0; c < 64; ++c) v += slots[(size_t)o.out * UC_SLOT_FLOATS + r * o.Cout + c] * P[pg.headWk + j * 64 + c];
          out[r * pg.head_dim + j] = v + P[pg.headB + j];
        }
  }
  if (q != pg.n_chunks) { fprintf(stderr, "chunk count mismatch %d vs %d\n", q, pg.n_chunks); return 1; }
  float err = 0.f;
  for (int i = 0; i < HD; ++i) err = fmaxf(err, fabsf(out[i] - expect[i]));
  printf("{\"max_abs_err\": %.6e, \"n_ops\": %d, \"n_chunks\": %d, \"stream_mb_per_cta\": %.3f, \"max_chunk_floats\": %d, \"smem_bytes\": %zu}\n", err,
         pg.n_ops, pg.n_chunks, S * 4 / 1e6, max_live_chunk_floats, uc_smem_bytes());
  return err <= 1e-4f ? 0 : 1;
}
