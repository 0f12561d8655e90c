// Image-encoder stem on the tensor cores (the caller in front of the sampling loop):
//   conv1 7x7 / stride 2 / pad 3 (3 -> 64) + folded BatchNorm + ReLU        modeling/resnet.py:193-197, 279-281   (stem_conv_kernel)
//   MaxPool2d(3, stride 2, pad 1)                                             modeling/resnet.py:198, 282           (maxpool_nhwc_kernel)
// for the encoder's bf16 mode.  cuDNN runs this layer (C_in = 3) as an Ampere-generation indexed implicit GEMM and torch's NHWC
// max-pool is 10x off the memory roofline: together 11.8 of the 22.3 ms the bf16 encoder spends on 256 camera frames
// (profiles/r02_encoder_profile.txt).  Here the convolution is a [pixels x 147] x [147 x 64] GEMM on tcgen05:
//   * a CTA owns tiles of 128 consecutive output pixels.  Its 256 threads gather the 7 x 7 x 3 input window of their pixel straight
//     from the fp32 image (any strides: NCHW or channels-last), round to bf16 and store it as the K-major, 128-byte-swizzled A
//     operand (K order = (kernel row, kernel column, channel), K padded 147 -> 160 = 10 instructions of K = 16);
//   * the folded weights arrive once per CTA as a pre-swizzled 24 KB image (bulk copy), B operand of every tile;
//   * accumulators in TMEM (64 fp32 columns); the epilogue adds the folded bias, applies ReLU and stores bf16 NHWC rows (128 B per pixel).
// Three CTAs per SM (72 KB of shared memory, 64 TMEM columns each) overlap one another's gather / MMA / store phases, so the kernel
// itself stays a simple sequential loop.  HBM-bound: 12 B read (once; the 12x window overlap is served by L1/L2) and 32 B written per
// input pixel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b2p {

constexpr int ST_THREADS = 256;
constexpr int ST_KREAL = 7 * 7 * 3;            // 147
constexpr int ST_UNITS = 19;                   // 16-byte units (8 bf16) holding real K elements
constexpr int ST_KSTEPS = 10;                  // K = 160: the unit after the last real one is zero, K steps 10 and 11 of the third chunk are never issued
constexpr int ST_A_CHUNK = 128 * 128;          // one 64-element K chunk of the A tile: [128 rows][128 B]
constexpr int ST_B_CHUNK = 64 * 128;           // one K chunk of the weight image: [64 channels][128 B]
constexpr int ST_A_BYTES = 3 * ST_A_CHUNK;
constexpr int ST_B_BYTES = 3 * ST_B_CHUNK;

struct __align__(16) StemShared {
  uint64_t wbar;       // weight image landed
  uint64_t mma_bar;    // the tile's MMAs have retired
  uint32_t tmem_base;
  uint32_t pad;
  float bias[64];
};

struct StemArgs {
  const float* img;
  long long sn;                 // element stride between images
  int sc, sh, sw;               // element strides inside an image
  int N, H, W, OH, OW;
  long long total;              // N * OH * OW output pixels
  int ntiles;
  const uint8_t* wimg;          // [3 chunks][64][128 B] bf16, swizzled (stem_weight_image() in modeling.py)
  const float* bias;            // [64] folded BatchNorm bias
  __nv_bfloat16* out;           // [N, OH, OW, 64]
};

size_t stem_smem_bytes() { return ST_A_BYTES + ST_B_BYTES + sizeof(StemShared) + 1024; }

__device__ __forceinline__ uint32_t st_swz(int r, int u) { return (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// units [J0, J1) of row m: 8 consecutive K elements each, K index = (r * 7 + kx) * 3 + c
template <int J0, int J1>
__device__ __forceinline__ void stem_gather(uint8_t* A, int m, const float* __restrict__ p, int off0, int sc, int sh, int sw, unsigned rowok, unsigned colok) {
#pragma unroll
  for (int j = J0; j < J1; ++j) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = j * 8 + e;
      if (k < ST_KREAL) {
        const int r = k / 21, kx = (k % 21) / 3, c = k % 3;
        const bool ok = ((rowok >> r) & 1u) && ((colok >> kx) & 1u);
        v[e] = ok ? __ldg(p + (off0 + r * sh + kx * sw + c * sc)) : 0.f;
      } else {
        v[e] = 0.f;
      }
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(A + (j >> 3) * ST_A_CHUNK + st_swz(m, j & 7)) = q;
  }
}

__global__ void __launch_bounds__(ST_THREADS, 3) stem_conv_kernel(const StemArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* A = smem;
  uint8_t* Bw = smem + ST_A_BYTES;
  StemShared* sh = reinterpret_cast<StemShared*>(smem + ST_A_BYTES + ST_B_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&sh->wbar, 1);
    mbar_init(&sh->mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&sh->wbar, ST_B_BYTES);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(Bw)), "l"(a.wimg),
                 "r"((uint32_t)ST_B_BYTES), "r"(smem_u32(&sh->wbar)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid >= 64 && tid < 128) sh->bias[tid - 64] = __ldg(a.bias + (tid - 64));
  if (tid < 128) *reinterpret_cast<uint4*>(A + 2 * ST_A_CHUNK + st_swz(tid, 3)) = make_uint4(0, 0, 0, 0);   // K 152..159: zero for every tile
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t idesc = umma_idesc_n(64);
  const int m = tid & 127, half = tid >> 7;
  const int opi = a.OH * a.OW;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    // ---- gather: this thread's half of the 7x7x3 window of output pixel g ----
    {
      const long long g = (long long)tile * 128 + m;
      unsigned rowok = 0, colok = 0;
      const float* p = a.img;
      int off0 = 0;
      if (g < a.total) {
        const int n = (int)(g / opi);
        const int rem = (int)(g - (long long)n * opi);
        const int oy = rem / a.OW, ox = rem - oy * a.OW;
        const int iy0 = 2 * oy - 3, ix0 = 2 * ox - 3;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          if (iy0 + r >= 0 && iy0 + r < a.H) rowok |= 1u << r;
          if (ix0 + r >= 0 && ix0 + r < a.W) colok |= 1u << r;
        }
        p = a.img + (long long)n * a.sn;
        off0 = iy0 * a.sh + ix0 * a.sw;
      }
      if (half == 0) stem_gather<0, 10>(A, m, p, off0, a.sc, a.sh, a.sw, rowok, colok);
      else stem_gather<10, ST_UNITS>(A, m, p, off0, a.sc, a.sh, a.sw, rowok, colok);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A -> visible to the tensor core
    __syncthreads();
    if (warp == 0 && elect_one()) {
      if (tile == (int)blockIdx.x) mbar_wait(&sh->wbar, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < ST_KSTEPS; ++ks) {
        const uint64_t ad = umma_desc(smem_u32(A) + (ks >> 2) * ST_A_CHUNK) + (uint64_t)((ks & 3) * 2);
        const uint64_t bd = umma_desc(smem_u32(Bw) + (ks >> 2) * ST_B_CHUNK) + (uint64_t)((ks & 3) * 2);
        umma(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(&sh->mma_bar);
    }
    mbar_wait(&sh->mma_bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- epilogue: thread = (row of its warp's TMEM lane quadrant, 32-channel half) ----
    {
      const int q = warp & 3, hf = warp >> 2;
      const int row = q * 32 + lane;
      float v[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);
      tmem_ld<16, false>(taddr, v);
      tmem_ld<16, false>(taddr + 16, v + 16);
      tmem_ld_wait();
      const long long g = (long long)tile * 128 + row;
      if (g < a.total) {
        uint4* dst = reinterpret_cast<uint4*>(a.out + g * 64 + hf * 32);
        const float* bs = sh->bias + hf * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16x2(fmaxf(v[8 * i + 0] + bs[8 * i + 0], 0.f), fmaxf(v[8 * i + 1] + bs[8 * i + 1], 0.f));
          o.y = pack_bf16x2(fmaxf(v[8 * i + 2] + bs[8 * i + 2], 0.f), fmaxf(v[8 * i + 3] + bs[8 * i + 3], 0.f));
          o.z = pack_bf16x2(fmaxf(v[8 * i + 4] + bs[8 * i + 4], 0.f), fmaxf(v[8 * i + 5] + bs[8 * i + 5], 0.f));
          o.w = pack_bf16x2(fmaxf(v[8 * i + 6] + bs[8 * i + 6], 0.f), fmaxf(v[8 * i + 7] + bs[8 * i + 7], 0.f));
          dst[i] = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM and the A tile are free for the next tile
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
  }
}

// MaxPool2d(3, 2, 1) on NHWC bf16: a thread owns 8 channels (16 bytes) of one output pixel; padding never wins (window centre is always inside)
__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W, int OH, int OW, int C8) {
  const long long total = (long long)N * OH * OW * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    long long px = i / C8;
    const int ox = (int)(px % OW); px /= OW;
    const int oy = (int)(px % OH);
    const int n = (int)(px / OH);
    const int y0 = 2 * oy - 1, x0 = 2 * ox - 1;
    const uint4* base = in + (long long)n * H * W * C8 + cg;
    uint4 best = __ldg(base + ((long long)(2 * oy) * W + 2 * ox) * C8);
    __nv_bfloat162* b2 = reinterpret_cast<__nv_bfloat162*>(&best);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = y0 + dy;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = x0 + dx;
        if (x < 0 || x >= W || (dy == 1 && dx == 1)) continue;
        uint4 v = __ldg(base + ((long long)y * W + x) * C8);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) b2[k] = __hmax2(b2[k], v2[k]);
      }
    }
    out[i] = best;
  }
}


// ------------------------------------------------------------------------------------------------------------------------------------
// conv1 + bn1 + relu + maxpool in ONE kernel (modeling/resnet.py:279-282).  The stand-alone pair above writes conv1's output (1.9 GB for
// 256 frames) and reads it back for the pool, and its gather of the 7x7x3 windows straight from global memory saturates the L1 data path
// (ncu: L1/TEX 80 % busy, 12x re-read of every input element).  Here a CTA sweeps down a strip of 30 + 2 conv columns of one image, four
// conv rows (= two pooled rows) per step:
//   * the 13-row x 69-column x 3-channel input patch of the step is staged in shared memory as fp32 [row][column][channel] (zeros outside the image), so the
//     window gather reads shared memory with compile-time offsets and no bounds checks;
//   * A operand / weights / tcgen05.mma exactly as in stem_conv_kernel (128 pixels x K = 160, N = 64);
//   * epilogue: + bias, ReLU, conv pixels outside conv1's output zeroed (the pool pads with -inf, but after the ReLU every window holds a
//     valid value >= 0, so 0 never wins wrongly), bf16 rows into a swizzled staging tile that reuses the A operand's memory;
//   * pool: 240 threads, one (pooled pixel, 8 channels) each: max over 3 x 3 staged conv pixels, the row above the step coming from a 4 KB
//     buffer that keeps the last conv row of the previous step; 16-byte stores, 128 contiguous bytes per pooled pixel.
// The pooled result is bit-identical to stem_conv_kernel + maxpool_nhwc_kernel (max of bf16-rounded values).
constexpr int SP_ROWS = 4, SP_COLS = 32;               // conv tile of a step: 4 rows x 32 columns = 128 pixels
constexpr int SP_PCOLS = 15;                           // pooled columns per strip (conv columns 30 s - 1 .. 30 s + 30)
constexpr int SP_PATCH_ROWS = 2 * SP_ROWS + 5;         // 13 input rows
constexpr int SP_PITCH = 208;                          // floats per staged patch row (69 * 3 = 207, padded: 8-byte aligned rows)
constexpr int SP_PATCH_BYTES = SP_PATCH_ROWS * SP_PITCH * 4;   // 10,816
constexpr int SP_PRE = (SP_PATCH_ROWS * 207 + ST_THREADS - 1) / ST_THREADS;   // patch values per thread (11)
constexpr int SP_KEEP_BYTES = SP_COLS * 128;           // last conv row of the previous step: [32 columns][64 ch] bf16

struct StemPoolArgs {
  const float* img;
  long long sn;
  int sc, sh, sw;
  int N, H, W, OH, OW, PH, PW;
  int strips, steps;            // ceil(PW / 15), ceil(PH / 2)
  int seg, nseg;                // steps per work item, items per strip: a strip is cut into row segments when there are few frames (each segment
                                // first recomputes the conv rows above it to fill the kept row: one extra step per segment)
  int n_items;                  // N * strips * nseg
  const uint8_t* wimg;
  const float* bias;
  __nv_bfloat16* out;           // [N, PH, PW, 64]
};

size_t stem_pool_smem_bytes() { return ST_A_BYTES + ST_B_BYTES + SP_PATCH_BYTES + SP_KEEP_BYTES + sizeof(StemShared) + 1024 + 64; }

// The 7x7x3 window of tile pixel (row, col) from the staged patch: K index k = (r * 7 + kx) * 3 + c -> patch[2 row + r][6 col + (k % 21)], i.e. 21 contiguous
// floats per kernel row.  Thread half 0 packs K units 0..10 (k < 88: kernel rows 0..3 and the first four values of row 4), half 1 units 11..18.  The floats
// are fetched as 8-byte pairs wherever (k % 21) is even and the pair stays inside the kernel row: lanes are consecutive columns = 24 B apart, which
// is conflict-free for 64-bit shared-memory loads (and a 2-way conflict for 32-bit ones).
template <int K0, int K1>
__device__ __forceinline__ void stem_gather_smem(uint8_t* A, int m, const float* __restrict__ p0) {
  float v[K1 - K0];
#pragma unroll
  for (int k = K0; k < K1; ++k) {                      // fully unrolled: every predicate below is a compile-time constant
    const int r = k / 21, q = k % 21;
    const bool first = k < ST_KREAL && (q & 1) == 0 && q + 1 < 21 && k + 1 < K1 && k + 1 < ST_KREAL;                       // first float of an 8-byte pair
    const bool second = k > K0 && k < ST_KREAL && ((k - 1) % 21 & 1) == 0 && (k - 1) % 21 + 1 < 21;                          // second float of the pair at k - 1
    if (k >= ST_KREAL) v[k - K0] = 0.f;
    else if (first) {
      const float2 t = *reinterpret_cast<const float2*>(p0 + r * SP_PITCH + q);
      v[k - K0] = t.x;
      v[k + 1 - K0] = t.y;
    } else if (!second) v[k - K0] = p0[r * SP_PITCH + q];
  }
#pragma unroll
  for (int j = K0 / 8; j < K1 / 8; ++j) {
    uint4 o;
    const float* u = v + (j * 8 - K0);
    o.x = pack_bf16x2(u[0], u[1]); o.y = pack_bf16x2(u[2], u[3]); o.z = pack_bf16x2(u[4], u[5]); o.w = pack_bf16x2(u[6], u[7]);
    *reinterpret_cast<uint4*>(A + (j >> 3) * ST_A_CHUNK + st_swz(m, j & 7)) = o;
  }
}

__global__ void __launch_bounds__(ST_THREADS, 2) stem_pool_kernel(const StemPoolArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* A = smem;                                   // also the conv staging tile of the epilogue: [128 pixels][128 B] swizzled
  uint8_t* Bw = smem + ST_A_BYTES;
  float* patch = reinterpret_cast<float*>(smem + ST_A_BYTES + ST_B_BYTES);
  uint8_t* keep = smem + ST_A_BYTES + ST_B_BYTES + SP_PATCH_BYTES;
  StemShared* sh = reinterpret_cast<StemShared*>(keep + SP_KEEP_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&sh->wbar, 1);
    mbar_init(&sh->mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&sh->wbar, ST_B_BYTES);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(Bw)), "l"(a.wimg),
                 "r"((uint32_t)ST_B_BYTES), "r"(smem_u32(&sh->wbar)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid >= 64 && tid < 128) sh->bias[tid - 64] = __ldg(a.bias + (tid - 64));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t idesc = umma_idesc_n(64);
  const int m = tid & 127, half = tid >> 7;
  const int trow = m >> 5, tcol = m & 31;              // tile pixel of this thread (gather) / TMEM lane
  uint32_t phase = 0;
  bool w_ready = false;

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int sg = item % a.nseg, ns = item / a.nseg;
    const int n = ns / a.strips, strip = ns - n * a.strips;
    const int st_begin = sg * a.seg, st_end = st_begin + a.seg < a.steps ? st_begin + a.seg : a.steps;
    const int cc0 = 2 * SP_PCOLS * strip - 1;          // first conv column of the strip
    const float* img = a.img + (long long)n * a.sn;
    *reinterpret_cast<uint4*>(keep + tid * 16) = make_uint4(0, 0, 0, 0);   // conv row -1: nothing (256 threads x 16 B = 4 KB)
    const int st_first = st_begin > 0 ? st_begin - 1 : 0;
    float pre[SP_PRE];
    auto load_patch = [&](int step, float (&v)[SP_PRE]) {
      const int iy0 = 2 * SP_ROWS * step - 3, ix0 = 2 * cc0 - 3;
#pragma unroll
      for (int j = 0; j < SP_PRE; ++j) {
        const int i = tid + j * ST_THREADS;
        const int pr = i / 207, rem = i - pr * 207;
        const int pc = rem / 3, c = rem - pc * 3;
        const int iy = iy0 + pr, ix = ix0 + pc;
        v[j] = (i < SP_PATCH_ROWS * 207 && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) ? __ldg(img + (long long)iy * a.sh + (long long)ix * a.sw + c * a.sc) : 0.f;
      }
    };
    for (int st = st_first; st < st_end; ++st) {
      const bool warm = st < st_begin;                 // the step above a segment: only its last conv row is wanted
      const int cr0 = SP_ROWS * st;                    // first conv row of the step
      // ---- 1. stage the input patch: rows 2 cr0 - 3 .., columns 2 cc0 - 3 ..  The values were fetched from global memory one step ahead (into
      //         registers, right after the previous step's patch was staged), so their latency hides behind that step's gather / MMA / epilogue / pool ----
      if (st == st_first) load_patch(st, pre);
#pragma unroll
      for (int j = 0; j < SP_PRE; ++j) {
        const int i = tid + j * ST_THREADS;
        if (i < SP_PATCH_ROWS * 207) patch[i + i / 207] = pre[j];      // row pitch 208 = 207 + 1
      }
      __syncthreads();
      if (st + 1 < st_end) load_patch(st + 1, pre);
      // ---- 2. A operand: this thread's half of the 7x7x3 window of tile pixel (trow, tcol) ----
      {
        const float* p0 = patch + (2 * trow) * SP_PITCH + 6 * tcol;
        if (half == 0) stem_gather_smem<0, 88>(A, m, p0);
        else stem_gather_smem<88, 8 * ST_UNITS>(A, m, p0);
        if (tid < 128) *reinterpret_cast<uint4*>(A + 2 * ST_A_CHUNK + st_swz(tid, 3)) = make_uint4(0, 0, 0, 0);   // K 152..159 (the staging tile overwrote it)
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      // ---- 3. MMAs ----
      if (warp == 0 && elect_one()) {
        if (!w_ready) mbar_wait(&sh->wbar, 0);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ST_KSTEPS; ++ks) {
          const uint64_t ad = umma_desc(smem_u32(A) + (ks >> 2) * ST_A_CHUNK) + (uint64_t)((ks & 3) * 2);
          const uint64_t bd = umma_desc(smem_u32(Bw) + (ks >> 2) * ST_B_CHUNK) + (uint64_t)((ks & 3) * 2);
          umma(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&sh->mma_bar);
      }
      w_ready = true;
      mbar_wait(&sh->mma_bar, phase);
      phase ^= 1u;
      tc_fence_after();
      // ---- 4. epilogue: thread = (tile pixel of its TMEM lane quadrant, 32-channel half) -> staging tile (reuses A: the MMAs have retired) ----
      {
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        float v[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);
        tmem_ld<16, false>(taddr, v);
        tmem_ld<16, false>(taddr + 16, v + 16);
        tmem_ld_wait();
        const int cr = cr0 + (row >> 5), cc = cc0 + (row & 31);
        const bool valid = cr < a.OH && cc >= 0 && cc < a.OW;
        const float* bs = sh->bias + hf * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (valid) {
            o.x = pack_bf16x2(fmaxf(v[8 * i + 0] + bs[8 * i + 0], 0.f), fmaxf(v[8 * i + 1] + bs[8 * i + 1], 0.f));
            o.y = pack_bf16x2(fmaxf(v[8 * i + 2] + bs[8 * i + 2], 0.f), fmaxf(v[8 * i + 3] + bs[8 * i + 3], 0.f));
            o.z = pack_bf16x2(fmaxf(v[8 * i + 4] + bs[8 * i + 4], 0.f), fmaxf(v[8 * i + 5] + bs[8 * i + 5], 0.f));
            o.w = pack_bf16x2(fmaxf(v[8 * i + 6] + bs[8 * i + 6], 0.f), fmaxf(v[8 * i + 7] + bs[8 * i + 7], 0.f));
          }
          *reinterpret_cast<uint4*>(A + st_swz(row, hf * 4 + i)) = o;
        }
      }
      tc_fence_before();
      __syncthreads();
      // ---- 5. pool: thread = (pooled row of the step, pooled column of the strip, 8 channels) ----
      if (!warm && tid < 2 * SP_PCOLS * 8) {
        const int prow = tid / (SP_PCOLS * 8), rem = tid - prow * (SP_PCOLS * 8);
        const int pcol = rem >> 3, c8 = rem & 7;
        const int pr = 2 * st + prow, pc = SP_PCOLS * strip + pcol;
        if (pr < a.PH && pc < a.PW) {
          uint4 best = make_uint4(0, 0, 0, 0);           // every staged value is >= 0
          __nv_bfloat162* b2 = reinterpret_cast<__nv_bfloat162*>(&best);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int lr = 2 * prow - 1 + dy;            // local conv row: -1 = the kept row
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int lc = 2 * pcol + dx;
              const uint4 val = lr < 0 ? *reinterpret_cast<const uint4*>(keep + lc * 128 + c8 * 16)
                                       : *reinterpret_cast<const uint4*>(A + st_swz(lr * 32 + lc, c8));
              const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
              for (int k = 0; k < 4; ++k) b2[k] = __hmax2(b2[k], v2[k]);
            }
          }
          *reinterpret_cast<uint4*>(a.out + (((long long)n * a.PH + pr) * a.PW + pc) * 64 + c8 * 8) = best;
        }
      }
      __syncthreads();
      // ---- 6. the step's last conv row is the next step's row -1 ----
      {
        const int lc = tid >> 3, c8 = tid & 7;
        *reinterpret_cast<uint4*>(keep + lc * 128 + c8 * 16) = *reinterpret_cast<const uint4*>(A + st_swz(3 * 32 + lc, c8));
      }
      __syncthreads();      // the staging tile (= A) may be overwritten by the next gather; keep is complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
  }
}

}  // namespace b2p

using namespace b2p;

extern "C" int b2p_encoder_stem_bf16(const float* img, int64_t stride_n, int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t N, int32_t H,
                                     int32_t W, const void* weight_image, const float* bias, void* out_nhwc_bf16, void* stream) {
  if (!img || !weight_image || !bias || !out_nhwc_bf16 || N <= 0 || H < 1 || W < 1) return B2P_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(weight_image) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc_bf16) & 15)) return B2P_ERR_INVALID_ARG;
  // offsets inside one image are 32-bit in the kernel
  const int64_t span = (stride_c < 0 ? -stride_c : stride_c) * 3 + (stride_h < 0 ? -stride_h : stride_h) * (int64_t)(H + 8) + (stride_w < 0 ? -stride_w : stride_w) * (int64_t)(W + 8);
  if (span >= (1LL << 31)) return B2P_ERR_INVALID_ARG;
  StemArgs a{};
  a.img = img; a.sn = stride_n; a.sc = (int)stride_c; a.sh = (int)stride_h; a.sw = (int)stride_w;
  a.N = N; a.H = H; a.W = W; a.OH = (H + 6 - 7) / 2 + 1; a.OW = (W + 6 - 7) / 2 + 1;
  a.total = (long long)N * a.OH * a.OW;
  const long long ntiles = (a.total + 127) / 128;
  if (ntiles >= (1LL << 31)) return B2P_ERR_INVALID_ARG;
  a.ntiles = (int)ntiles;
  a.wimg = reinterpret_cast<const uint8_t*>(weight_image); a.bias = bias; a.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc_bf16);
  const int smem = (int)stem_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    B2P_CUDA_TRY(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    B2P_CUDA_TRY(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  const int grid = a.ntiles < 148 * 3 ? a.ntiles : 148 * 3;
  stem_conv_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int b2p_maxpool3x3s2_nhwc_bf16(const void* in, void* out, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!in || !out || N <= 0 || H < 1 || W < 1 || C < 8 || C % 8) return B2P_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return B2P_ERR_INVALID_ARG;
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = (long long)N * OH * OW * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 148LL * 8 * 4;
  if (blocks > cap) blocks = cap;
  maxpool_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), N, H, W, OH, OW, C / 8);
  return (int)cudaGetLastError();
}

extern "C" int b2p_encoder_stem_pool_bf16(const float* img, int64_t stride_n, int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t N, int32_t H,
                                          int32_t W, const void* weight_image, const float* bias, void* out_nhwc_bf16, void* stream) {
  if (!img || !weight_image || !bias || !out_nhwc_bf16 || N <= 0 || H < 1 || W < 1) return B2P_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(weight_image) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc_bf16) & 15)) return B2P_ERR_INVALID_ARG;
  if (stride_c < 0 || stride_h < 0 || stride_w < 0) return B2P_ERR_INVALID_ARG;
  const int64_t span = stride_c * 3 + stride_h * (int64_t)(H + 8) + stride_w * (int64_t)(W + 8);
  if (span >= (1LL << 31)) return B2P_ERR_INVALID_ARG;
  StemPoolArgs a{};
  a.img = img; a.sn = stride_n; a.sc = (int)stride_c; a.sh = (int)stride_h; a.sw = (int)stride_w;
  a.N = N; a.H = H; a.W = W; a.OH = (H - 1) / 2 + 1; a.OW = (W - 1) / 2 + 1; a.PH = (a.OH - 1) / 2 + 1; a.PW = (a.OW - 1) / 2 + 1;
  a.strips = (a.PW + SP_PCOLS - 1) / SP_PCOLS; a.steps = (a.PH + 1) / 2;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  a.seg = a.steps;
  while (a.seg > 4 && (long long)N * a.strips * ((a.steps + a.seg - 1) / a.seg) < 2LL * sms) a.seg = (a.seg + 1) / 2;   // few frames: shorter segments fill the chip
  a.nseg = (a.steps + a.seg - 1) / a.seg;
  const long long items = (long long)N * a.strips * a.nseg;
  if (items >= (1LL << 31)) return B2P_ERR_INVALID_ARG;
  a.n_items = (int)items;
  a.wimg = reinterpret_cast<const uint8_t*>(weight_image); a.bias = bias; a.out = reinterpret_cast<__nv_bfloat16*>(out_nhwc_bf16);
  const int smem = (int)stem_pool_smem_bytes();
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    B2P_CUDA_TRY(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    B2P_CUDA_TRY(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int grid = a.n_items < sms * 2 ? a.n_items : sms * 2;
  stem_pool_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(a);
  return (int)cudaGetLastError();
}
