// Image-encoder body on the tensor cores: every convolution of ResNet-34's layer1..layer4 with its folded BatchNorm,
// the residual add and the ReLU                                          modeling/resnet.py:56-102 (BasicBlock), 199-214, 283-286
//   3x3 / stride 1 / pad 1,   3x3 / stride 2 / pad 1,   1x1 / stride 2 (projection shortcut)      on NHWC bf16 activations, fp32 accumulation.
//
// Implicit GEMM without im2col AND without re-reading the input per tap:
//   * an output tile is 16 rows x 8 columns of pixels of one image (M = 128).  Its input for one 64-channel K chunk is ONE halo patch of
//     18 x 10 pixels, a single 4-D TMA box (out-of-bounds rows / columns / images arrive as zeros = the convolution's padding);
//   * the patch lands as [pixel][128 B] with the 128-byte swizzle, which is a function of the shared-memory ADDRESS — so the A operand of
//     tap (dy, dx) is the same patch read through a descriptor that starts (dy * 10 + dx) * 128 B later, with 8-row groups one patch row
//     (1280 B) apart (scripts/conv3x3_probe.cu: exact for all nine taps with the descriptor's base-offset field left 0).  The input is
//     read from L2 once per tile and chunk instead of nine times;
//   * stride 2: the input is viewed as four parity sub-images (even / odd rows x columns), each a 4-D tensor map with doubled strides; a tap
//     is then a unit-stride window of one parity patch, so the same machinery applies with four patches per tile and chunk;
//   * weights [tap][C_out][C_in] stream through a ring of (tap, chunk) stages of NB output channels; a work item is G tiles x NB channels, so
//     a weight stage is used by G tiles (G x 4 instructions of N = NB), accumulators in TMEM (G x NB columns; two sets when they fit, so
//     the epilogue of one item overlaps the MMAs of the next);
//   * persistent CTAs (one per SM) walk the items channel-block-major, so the CTAs running at the same time share their weight stages in L2;
//   * warp roles: 0 = patch producer, 2 = weight producer, 1 = MMA issuer (elected lanes), 4..7 = epilogue (TMEM -> + bias (+ residual)
//     -> ReLU -> bf16 -> NHWC rows, 16-byte stores).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b2p {

constexpr int EC_THREADS = 384;                     // warps 0..2: producers / MMA issuer, 3: idle, 4..11: epilogue (two per TMEM lane quadrant)
constexpr int EC_PW = 10, EC_PH = 18;                 // halo patch: (8 + 2) x (16 + 2) pixels
constexpr int EC_PATCH = EC_PW * EC_PH * 128;         // 23,040 B written by one TMA box
constexpr int EC_ASLOT = 23 * 1024;                   // patch slot (1024-aligned)
constexpr int EC_MAX_A = 8, EC_MAX_W = 4;
constexpr int EC_STG = 128 * 128;                     // one staged slab: 128 pixels x 64 channels bf16

struct EncMaps {
  CUtensorMap a[4];     // input: one map (stride 1) or the four parity sub-images (stride 2), dims (C, W, H, N)
  CUtensorMap w;        // weights (C_in, C_out, taps)
  CUtensorMap out, res; // output / residual (C_out, W, H, N): 64 channels x one tile per box (the epilogue's TMA store / load)
};

struct EncArgs {
  int N, H, W;                  // OUTPUT pixels
  int Cin, Cout, NB, G;
  int tiles_x, tiles_y, n_tiles;
  int n_groups, n_items;        // n_items = n_groups * (Cout / NB), channel-block major
  int n_maps, n_taps;
  int map_dx[4], map_dy[4];     // patch origin of map m relative to the tile origin (in that map's pixel grid)
  int tap_map[9], tap_off[9], tap_w[9];   // patch of the tap, byte offset of its window inside the patch, weight-map tap coordinate
  int na, nw, sets;             // patch slots, weight slots, accumulator sets
  int nstg;                     // epilogue staging buffers (16 KB each, for the output and for the residual): 1 or 2
  int relu;
  int transposed;               // 0: tile = 16 rows x 8 columns, patch stored [row][column];  1: tile = 8 rows x 16 columns, patch stored [column][row]
                                // (the input map lists H before W): for feature maps of height <= 8, where a 16-row tile would be half empty
  const float* bias;
  const __nv_bfloat16* res;     // NHWC, same shape as out (may be null)
  __nv_bfloat16* out;
};

struct __align__(16) EncShared {
  uint64_t a_full[EC_MAX_A], a_empty[EC_MAX_A];
  uint64_t w_full[EC_MAX_W], w_empty[EC_MAX_W];
  uint64_t t_full[2], t_empty[2];
  uint64_t res_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* src, const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t umma_desc_sbo(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(EC_THREADS, 1) enc_conv_kernel(const __grid_constant__ EncMaps maps, const EncArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  const int wslot_bytes = a.NB * 128;
  uint8_t* sW = smem + a.na * EC_ASLOT;
  uint8_t* sOut = sW + a.nw * wslot_bytes;            // nstg x 16 KB: a tile's 64-channel slab on its way out (swizzled, TMA store)
  uint8_t* sRes = sOut + a.nstg * EC_STG;             // nstg x 16 KB: the residual slab on its way in (TMA load)
  EncShared* sh = reinterpret_cast<EncShared*>(sRes + a.nstg * EC_STG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < a.na; ++i) { mbar_init(&sh->a_full[i], 1); mbar_init(&sh->a_empty[i], 1); }
    for (int i = 0; i < a.nw; ++i) { mbar_init(&sh->w_full[i], 1); mbar_init(&sh->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh->t_full[i], 1); mbar_init(&sh->t_empty[i], 8); mbar_init(&sh->res_full[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();                   // the next convolution may set up (barriers, TMEM, first weight stages) as SMs free up
  const uint32_t tmem_base = sh->tmem_base;
  const int chunks = a.Cin >> 6;
  const int ppc = a.G * a.n_maps;                  // patches per K chunk of an item
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      // =============================== patch producer ===============================
      for (int m = 0; m < a.n_maps; ++m) prefetch_tmap(&maps.a[m]);
      griddep_wait();                            // the input is the previous kernel's output
      int slot = 0; uint32_t par = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int gi = item % a.n_groups;
        for (int ch = 0; ch < chunks; ++ch) {
          for (int g = 0; g < a.G; ++g) {
            const int tile = gi * a.G + g;
            int n = a.N, y0 = 0, x0 = 0;           // a tile past the end reads image N: all zeros
            if (tile < a.n_tiles) {
              n = tile / tiles_per_img;
              const int rem = tile - n * tiles_per_img;
              const int ty = rem / a.tiles_x;
              y0 = ty * (a.transposed ? 8 : 16); x0 = (rem - ty * a.tiles_x) * (a.transposed ? 16 : 8);
            }
            for (int m = 0; m < a.n_maps; ++m) {
              mbar_wait(&sh->a_empty[slot], par ^ 1u);
              mbar_expect_tx(&sh->a_full[slot], EC_PATCH);
              if (a.transposed) tma_load_4d(sA + slot * EC_ASLOT, &maps.a[m], &sh->a_full[slot], ch * 64, y0 + a.map_dy[m], x0 + a.map_dx[m], n);
              else tma_load_4d(sA + slot * EC_ASLOT, &maps.a[m], &sh->a_full[slot], ch * 64, x0 + a.map_dx[m], y0 + a.map_dy[m], n);
              if (++slot == a.na) { slot = 0; par ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      // =============================== weight producer ===============================
      prefetch_tmap(&maps.w);
      int slot = 0; uint32_t par = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int nb = item / a.n_groups;
        for (int ch = 0; ch < chunks; ++ch) {
          for (int t = 0; t < a.n_taps; ++t) {
            mbar_wait(&sh->w_empty[slot], par ^ 1u);
            mbar_expect_tx(&sh->w_full[slot], (uint32_t)wslot_bytes);
            tma_load_3d(sW + slot * wslot_bytes, &maps.w, &sh->w_full[slot], ch * 64, nb * a.NB, a.tap_w[t]);
            if (++slot == a.nw) { slot = 0; par ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // =============================== MMA issuer ===============================
      const uint32_t idesc = umma_idesc_n(a.NB);
      const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
      int aslot = 0, wslot = 0, it = 0;
      uint32_t apar = 0, wpar = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        const int set = it & (a.sets - 1);
        const uint32_t tpar = (uint32_t)((it / a.sets) & 1);
        mbar_wait(&sh->t_empty[set], tpar ^ 1u);   // the epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t dbase = tmem_base + (uint32_t)(set * a.G * a.NB);
        for (int ch = 0; ch < chunks; ++ch) {
          for (int p = 0; p < ppc; ++p) mbar_wait(&sh->a_full[aslot + p], apar);
          tc_fence_after();
          for (int t = 0; t < a.n_taps; ++t) {
            mbar_wait(&sh->w_full[wslot], wpar);
            tc_fence_after();
            const uint64_t bd = umma_desc_sbo(sW_addr + wslot * wslot_bytes, 1024);
            for (int g = 0; g < a.G; ++g) {
              const uint64_t ad = umma_desc_sbo(sA_addr + (aslot + g * a.n_maps + a.tap_map[t]) * EC_ASLOT + a.tap_off[t], EC_PW * 128);
              const uint32_t d = dbase + (uint32_t)(g * a.NB);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma(d, ad + 2 * k, bd + 2 * k, idesc, (ch | t | k) ? 1u : 0u);
            }
            umma_commit(&sh->w_empty[wslot]);
            if (++wslot == a.nw) { wslot = 0; wpar ^= 1u; }
          }
          for (int p = 0; p < ppc; ++p) umma_commit(&sh->a_empty[aslot + p]);
          aslot += ppc;
          if (aslot == a.na) { aslot = 0; apar ^= 1u; }
        }
        umma_commit(&sh->t_full[set]);
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue (warps 4..11) ===============================
    // An item is G tiles x NB / 64 slabs of 128 pixels x 64 channels.  A thread owns one pixel (its TMEM lane) and 32 of the slab's channels
    // (warps 4..7: the lower half, 8..11: the upper half).  Neither the residual nor the result is touched with per-thread global accesses
    // (a warp would hit 32 different 128-byte lines per instruction: that made the first version LSU-bound): the residual slab arrives by TMA
    // into a swizzled staging buffer two slabs ahead, the result is written into a second swizzled buffer and leaves by a TMA store, which
    // also clips ragged tiles.
    const int q = warp & 3, e = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int th = a.transposed ? 8 : 16, tw = a.transposed ? 16 : 8;
    const int spt = a.NB >> 6, nsl = a.G * spt;                    // slabs per tile / per item
    const bool leader = warp == 4 && lane == 0;
    griddep_wait();                              // residual reads and output writes are ordered behind the previous kernel
    const uint32_t swz = (uint32_t)(m & 7);
    uint32_t rpar = 0u;                                            // bit b: phase parity of res_full[b]
    int it = 0;
    // origin of slab ls of the current item: false when its tile lies past the last image (warp-uniform)
    auto slab_origin = [&](int gi, int nb, int ls, int* cc, int* x0, int* y0, int* n) -> bool {
      const int g = ls / spt, sb = ls - g * spt;
      const int tile = gi * a.G + g;
      if (tile >= a.n_tiles) return false;
      *n = tile / tiles_per_img;
      const int rem = tile - *n * tiles_per_img;
      const int ty = rem / a.tiles_x;
      *y0 = ty * th; *x0 = (rem - ty * a.tiles_x) * tw;
      *cc = nb * a.NB + sb * 64;
      return true;
    };
    auto load_res = [&](int gi, int nb, int ls, int buf) {         // leader only
      int cc, x0, y0, n;
      if (ls < nsl && slab_origin(gi, nb, ls, &cc, &x0, &y0, &n)) {
        mbar_expect_tx(&sh->res_full[buf], EC_STG);
        if (a.transposed) tma_load_4d(sRes + buf * EC_STG, &maps.res, &sh->res_full[buf], cc, y0, x0, n);
        else tma_load_4d(sRes + buf * EC_STG, &maps.res, &sh->res_full[buf], cc, x0, y0, n);
      }
    };
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      const int set = it & (a.sets - 1);
      const uint32_t tpar = (uint32_t)((it / a.sets) & 1);
      const int nb = item / a.n_groups, gi = item - nb * a.n_groups;
      const float* bias = a.bias + nb * a.NB;
      if (leader && a.res) {                                        // the first residual slabs travel while the MMAs of the item still run
        load_res(gi, nb, 0, 0);
        if (a.nstg == 2) load_res(gi, nb, 1, 1);
      }
      mbar_wait_sleep(&sh->t_full[set], tpar);
      tc_fence_after();
      for (int ls = 0; ls < nsl; ++ls) {
        int cc, x0, y0, n;
        if (!slab_origin(gi, nb, ls, &cc, &x0, &y0, &n)) break;
        const int buf = a.nstg == 2 ? (ls & 1) : 0;
        const int g = ls / spt, sb = ls - g * spt;
        float v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * a.G * a.NB + g * a.NB + sb * 64 + e * 32);
        tmem_ld<16, false>(taddr, v);
        tmem_ld<16, false>(taddr + 16, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + sb * 64 + e * 32 + i));
          v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
        }
        if (a.res) {
          mbar_wait(&sh->res_full[buf], (rpar >> buf) & 1u);
          rpar ^= 1u << buf;
          const uint8_t* rrow = sRes + buf * EC_STG + m * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rrow + (((uint32_t)(e * 4 + j) ^ swz) << 4));
            const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(&rv);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[8 * j + k] += __bfloat162float(rb[k]);
          }
        }
        if (a.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        uint8_t* orow = sOut + buf * EC_STG + m * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack2(v[8 * j], v[8 * j + 1]); o.y = pack2(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack2(v[8 * j + 4], v[8 * j + 5]); o.w = pack2(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(orow + (((uint32_t)(e * 4 + j) ^ swz) << 4)) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the TMA store
        epi_bar();                                                      // A: the slab is staged, the residual buffer has been read
        if (leader) {
          if (a.transposed) tma_store_4d(sOut + buf * EC_STG, &maps.out, cc, y0, x0, n);
          else tma_store_4d(sOut + buf * EC_STG, &maps.out, cc, x0, y0, n);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (a.res) load_res(gi, nb, ls + a.nstg, buf);
          // the buffer the NEXT slab writes must have been read by its store: with two buffers that is the store before this one
          if (a.nstg == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        epi_bar();                                                      // B: the other staging buffer is free
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->t_empty[set]);
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores have been written before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace b2p

using namespace b2p;

extern "C" int b2p_encoder_conv_bf16(const void* in_nhwc, int32_t N, int32_t H, int32_t W, int32_t Cin, const void* w_packed, const float* bias, const void* res_nhwc,
                                     void* out_nhwc, int32_t Cout, int32_t ksize, int32_t stride, int32_t relu, void* stream) {
  if (!in_nhwc || !w_packed || !bias || !out_nhwc || N <= 0 || H < 1 || W < 1) return B2P_ERR_INVALID_ARG;
  if (Cin < 64 || Cin % 64 || Cout < 64 || Cout % 64) return B2P_ERR_INVALID_ARG;
  if (!((ksize == 3 && (stride == 1 || stride == 2)) || (ksize == 1 && stride == 2))) return B2P_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(in_nhwc) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(out_nhwc) | reinterpret_cast<uintptr_t>(res_nhwc)) & 15) return B2P_ERR_INVALID_ARG;
  EncodeTiledFn enc = enc_get_encode();
  if (!enc) return B2P_ERR_NO_DEVICE;
  EncArgs a;
  EncMaps m;
  memset(&a, 0, sizeof(a));
  memset(&m, 0, sizeof(m));
  a.N = N; a.H = (H - 1) / stride + 1; a.W = (W - 1) / stride + 1;
  a.Cin = Cin; a.Cout = Cout;
  a.relu = relu; a.bias = bias; a.res = reinterpret_cast<const __nv_bfloat16*>(res_nhwc); a.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc);
  a.transposed = (stride == 1 && a.H <= 8) ? 1 : 0;
  a.tiles_x = a.transposed ? (a.W + 15) / 16 : (a.W + 7) / 8; a.tiles_y = a.transposed ? (a.H + 7) / 8 : (a.H + 15) / 16;
  const long long nt = (long long)N * a.tiles_x * a.tiles_y;
  if (nt >= (1LL << 30)) return B2P_ERR_INVALID_ARG;
  a.n_tiles = (int)nt;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  a.n_maps = (ksize == 3 && stride == 2) ? 4 : 1;
  // tiles per item (G), channel-block width (NB), accumulator sets, ring depths.  Shared memory: na patch slots of 23 KB + nw weight slots of
  // NB * 128 B + 2 nstg staging slabs of 16 KB <= ~224 KB; TMEM: sets * G * NB <= 512 columns.
  // (measured on 256 frames, 64 -> 64: G = 1: 0.62 ms, G = 2: 0.44, G = 4: 0.44, one accumulator set: 0.60, the 72 KB filter resident with a
  //  3-patch ring: 0.54 — what counts for that HBM-bound layer is the input bytes in flight and the overlapped epilogue)
  auto configure = [&](int nb, int g) {
    a.NB = nb; a.G = g; a.na = 4;
    if (nb == 256) { a.sets = g == 1 ? 2 : 1; a.nw = 3; a.nstg = 1; }      // 94 + 96 + 32 KB
    else { a.sets = 2; a.nw = 4; a.nstg = 2; }                              // 94 + 32 / 64 + 64 KB
    a.n_groups = (a.n_tiles + a.G - 1) / a.G;
    return (long long)a.n_groups * (Cout / a.NB);
  };
  // widest channel block: 128 with two tiles per item (256 TMEM columns, two sets: the epilogue overlaps the next item's MMAs; 256 -> 256: 0.26 -> 0.24 ms
  // against 256-wide blocks with one set) — except for the 3x3/2 layers, whose single-tile items would stream the weights twice as often (0.26 -> 0.41 ms)
  const int nbmax = a.n_maps == 1 ? 128 : 256;
  int nb0 = nbmax;
  while (Cout % nb0) nb0 >>= 1;                    // widest allowed block that divides C_out (64 always does)
  long long items = configure(nb0, a.n_maps == 1 ? 2 : 1);
  // few frames (a closed-loop tick encodes ONE): the default items would leave most SMs idle and each CTA with a long serial K loop, so
  // split finer — one tile per item first, then narrower channel blocks — until the launch covers the chip
  while (items < sms && (a.G > 1 || a.NB > 64)) items = a.G > 1 ? configure(a.NB, 1) : configure(a.NB / 2, 1);
  const char* base = reinterpret_cast<const char*>(in_nhwc);
  const cuuint32_t box[4] = {64, EC_PW, EC_PH, 1}, estr[4] = {1, 1, 1, 1};
  if (stride == 1) {
    a.n_taps = 9; a.map_dx[0] = -1; a.map_dy[0] = -1;
    for (int t = 0; t < 9; ++t) { a.tap_map[t] = 0; a.tap_off[t] = (a.transposed ? (t % 3) * EC_PW + t / 3 : (t / 3) * EC_PW + t % 3) * 128; a.tap_w[t] = t; }
    // transposed: the map lists H before W, so the {64, 10, 18} box lands [column][row][channel]
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)(a.transposed ? H : W), (cuuint64_t)(a.transposed ? W : H), (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)(a.transposed ? W : 1) * Cin * 2, (cuuint64_t)(a.transposed ? 1 : W) * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    if (enc(&m.a[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B2P_ERR_INVALID_ARG;
  } else {
    // parity sub-images: map (py, px) holds the input pixels (2 yy + py, 2 xx + px)
    const int np = a.n_maps;
    for (int p = 0; p < np; ++p) {
      const int py = p >> 1, px = p & 1;
      const int Hp = (H - py + 1) / 2, Wp = (W - px + 1) / 2;
      a.map_dx[p] = px ? -1 : 0; a.map_dy[p] = py ? -1 : 0;
      cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)(Wp > 0 ? Wp : 1), (cuuint64_t)(Hp > 0 ? Hp : 1), (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 2};
      if (Wp <= 0 || Hp <= 0) return B2P_ERR_INVALID_ARG;
      if (enc(&m.a[p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base) + ((size_t)py * W + px) * Cin * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B2P_ERR_INVALID_ARG;
    }
    if (ksize == 3) {
      a.n_taps = 9;
      for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t % 3;
        const int py = dy != 1, px = dx != 1, oy = dy == 2, ox = dx == 2;   // input row 2 y + dy - 1: odd for dy = 0 (yy = y - 1) and dy = 2 (yy = y)
        a.tap_map[t] = py * 2 + px; a.tap_off[t] = (oy * EC_PW + ox) * 128; a.tap_w[t] = t;
      }
    } else {
      a.n_taps = 1; a.tap_map[0] = 0; a.tap_off[0] = 0; a.tap_w[0] = 0;
    }
  }
  {
    const int taps = ksize * ksize;
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t wbox[3] = {64, (cuuint32_t)a.NB, 1}, westr[3] = {1, 1, 1};
    if (enc(&m.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_packed), dims, strides, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B2P_ERR_INVALID_ARG;
  }
  {  // output / residual: one 64-channel slab of a tile per box, same tile orientation as the input patches
    const int OH = a.H, OW = a.W;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)(a.transposed ? OH : OW), (cuuint64_t)(a.transposed ? OW : OH), (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)(a.transposed ? OW : 1) * Cout * 2, (cuuint64_t)(a.transposed ? 1 : OW) * Cout * 2, (cuuint64_t)OH * OW * Cout * 2};
    cuuint32_t obox[4] = {64, 8, 16, 1}, oestr[4] = {1, 1, 1, 1};
    if (enc(&m.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out_nhwc, dims, strides, obox, oestr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B2P_ERR_INVALID_ARG;
    if (res_nhwc && enc(&m.res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(res_nhwc), dims, strides, obox, oestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return B2P_ERR_INVALID_ARG;
  }
  if (items >= (1LL << 31)) return B2P_ERR_INVALID_ARG;
  a.n_items = (int)items;
  const size_t smem = (size_t)a.na * EC_ASLOT + (size_t)a.nw * a.NB * 128 + (size_t)2 * a.nstg * EC_STG + sizeof(EncShared) + 1024;
  if (smem > 227 * 1024 || a.na > EC_MAX_A || a.nw > EC_MAX_W || a.na % (a.G * a.n_maps)) return B2P_ERR_INVALID_ARG;
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {     // function attributes are per device
    B2P_CUDA_TRY(cudaFuncSetAttribute(enc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2P_CUDA_TRY(cudaFuncSetAttribute(enc_conv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int grid = a.n_items < sms ? a.n_items : sms;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(EC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // weights do not depend on the previous layer: their first stages overlap its tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, enc_conv_kernel, m, a);
}
