// Camera frame -> encoder input on the device.  Replaces torchvision's ToTensor + Normalize
// (interact.py:72-77, 170-172) for uint8 frames that were uploaded as bytes: out = ((float(u8) / 255) - mean[c]) / std[c],
// same operations in the same order, so the result is bit-identical to the host-side transform.  Output order == input
// order (N,H,W,C): that IS the channels-last memory of the logical [N,3,H,W] tensor the encoder consumes.
// HBM-bound byte work: 3 B read + 12 B written per pixel.  A thread converts 4 consecutive ELEMENTS (one 32-bit load, one
// 128-bit store), so a warp reads 128 and writes 512 contiguous bytes; the channel of element e is e % 3.
#include "common.cuh"

namespace b2p {

struct Norm3 { float mean[3], std[3]; };

__device__ __forceinline__ float norm1(uint32_t byte, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)byte, 255.0f), mean), std);
}

// Only 3 x 256 distinct results exist: every CTA first builds them with the exact divisions in shared memory, then the
// streaming loop is one 32-bit load, four table reads and one 128-bit store per thread.
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long n_elems, Norm3 nm) {
  __shared__ float lut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int c = i >> 8;
    lut[i] = norm1((uint32_t)(i & 255), c == 0 ? nm.mean[0] : (c == 1 ? nm.mean[1] : nm.mean[2]), c == 0 ? nm.std[0] : (c == 1 ? nm.std[1] : nm.std[2]));
  }
  __syncthreads();
  const long long groups = n_elems >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uint32_t* in32 = reinterpret_cast<const uint32_t*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint32_t w = __ldg(in32 + g);
    const int c0 = (int)(g % 3);                                  // (4 g) % 3 == g % 3
    const int c1 = c0 == 2 ? 0 : c0 + 1, c2 = c1 == 2 ? 0 : c1 + 1;
    float4 v;
    v.x = lut[(c0 << 8) + (w & 0xffu)];
    v.y = lut[(c1 << 8) + ((w >> 8) & 0xffu)];
    v.z = lut[(c2 << 8) + ((w >> 16) & 0xffu)];
    v.w = lut[(c0 << 8) + (w >> 24)];
    out4[g] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)                        // tail elements (n_elems % 4)
    for (long long e = groups << 2; e < n_elems; ++e) out[e] = lut[((int)(e % 3) << 8) + in[e]];
}

}  // namespace b2p

extern "C" int b2p_preprocess_frames(const uint8_t* frames_nhwc, float* out_nhwc, int64_t n_pixels, const float mean[3], const float std_[3],
                                     void* stream) {
  if (n_pixels < 0 || (n_pixels > 0 && (!frames_nhwc || !out_nhwc)) || !mean || !std_) return B2P_ERR_INVALID_ARG;
  if (n_pixels == 0) return B2P_OK;
  if ((reinterpret_cast<uintptr_t>(frames_nhwc) & 3) || (reinterpret_cast<uintptr_t>(out_nhwc) & 15)) return B2P_ERR_INVALID_ARG;
  b2p::Norm3 nm;
  for (int c = 0; c < 3; ++c) {
    if (!(std_[c] != 0.f)) return B2P_ERR_INVALID_ARG;
    nm.mean[c] = mean[c]; nm.std[c] = std_[c];
  }
  const long long n_elems = (long long)n_pixels * 3;
  const long long groups = n_elems >> 2;
  long long blocks = (groups + 255) / 256;
  const long long cap = 148LL * 16;                              // resident CTAs of 256 threads on 148 SMs, grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  b2p::preprocess_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frames_nhwc, out_nhwc, n_elems, nm);
  return (int)cudaGetLastError();
}
