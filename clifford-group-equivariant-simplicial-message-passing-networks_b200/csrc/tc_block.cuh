// Shared definitions of the tensor-core CEMLP-block kernels (tc_block_fwd.cu, tc_block_bwd.cu).
//
// BPT ("blade-plane tile") tensor layout, used for every intermediate of the tensor-core path:
//     [tile = row / 128][blade][channel / 4][row % 128][channel % 4]      fp32, channels padded to Cp (multiple of 16)
// * one (tile, blade, 8-channel chunk) is 4 KB contiguous and already in the UMMA no-swizzle K-major operand layout,
//   so a GEMM kernel brings it in with two bulk copies and no transposition;
// * a thread that owns one row reads/writes 4 channels of a blade as one float4, and consecutive lanes (rows) touch
//   consecutive 16-byte units: every epilogue access is perfectly coalesced.
// Padded rows and padded channels always hold zeros.
#pragma once
#include "tc_common.cuh"

namespace csmpn {
namespace tcb {
using namespace tc;

constexpr int kTile = 128;     // rows per tile = UMMA M
constexpr int kThreads = 512;  // 16 warps: warp w owns TMEM lanes 32*(w%4).., channel groups (w/4), (w/4)+4, ...
constexpr float kInvSqrt2 = 0.70710678118654752440f;

// chunk buffer (one K step of 8 channels, all blades): hi planes then lo planes.  A plane is [2][128][4] fp32; the two
// halves are KH apart and planes PS apart, padded so that the scalar stores of the transposing producer (lanes differ
// in channel and blade half) spread over all 32 banks.
constexpr uint32_t kKH = 2048 + 32;
constexpr uint32_t kPS = 2 * kKH + 16;

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int64_t bpt_floats(int B, int64_t rows, int cp) {
  return ((rows + kTile - 1) / kTile) * (int64_t)B * cp * kTile;
}
// float offset of (tile, blade, channel group c4, row r) -> 4 consecutive channels
__device__ __forceinline__ size_t bpt_off(int B, int cp, int64_t tile, int b, int c4, int r) {
  return ((((size_t)tile * B + b) * (cp >> 2) + c4) * kTile + r) * 4;
}

template <int DIM>
__device__ __forceinline__ void silu_gates(const float* y1, const float* a, const float* b, float* sg, float* inv) {
  using A = Alg<DIM>;
#pragma unroll
  for (int g = 0; g < A::G; ++g) inv[g] = 0.f;
#pragma unroll
  for (int i = 0; i < A::B; ++i) inv[A::grade_of(i)] = fmaf(y1[i], y1[i], inv[A::grade_of(i)]);
  inv[0] = y1[0];
#pragma unroll
  for (int g = 0; g < A::G; ++g) sg[g] = sigmoidf_(fmaf(a[g], inv[g], b[g]));
}

// normalisation: xn_i = xr_i * rinv[g];  den_g = s_g (nrm_g - 1) + 1 + eps
template <int DIM>
__device__ __forceinline__ void norm_factors(const float* xr, const float* s, float* q, float* nrm, float* rinv) {
  using A = Alg<DIM>;
#pragma unroll
  for (int g = 0; g < A::G; ++g) q[g] = 0.f;
#pragma unroll
  for (int i = 0; i < A::B; ++i) q[A::grade_of(i)] = fmaf(xr[i], xr[i], q[A::grade_of(i)]);
#pragma unroll
  for (int g = 0; g < A::G; ++g) {
    nrm[g] = smooth_abs_sqrt(q[g]);
    rinv[g] = 1.f / (fmaf(s[g], nrm[g] - 1.f, 1.f) + kEps);
  }
}

template <int DIM>
__device__ __forceinline__ float mv_sumsq(const float* x) {
  float Q = 0.f;
#pragma unroll
  for (int i = 0; i < Alg<DIM>::B; ++i) Q = fmaf(x[i], x[i], Q);
  return Q;
}

// Stage a [c_out, c_in, G] weight (global, reference layout) as 2*G operand images in shared memory:
//   image (g, hl) = plane with R = rows_p rows;  TRANS == false: row = output channel, column = input channel
//                                                TRANS == true : row = input channel,  column = output channel
// hl = 0: TF32-exact high part, hl = 1: remainder.  Rows/columns beyond the real sizes are zero.
template <int DIM, bool TRANS>
__device__ __forceinline__ void stage_weight_images(uint8_t* img0, uint32_t img_bytes, const float* __restrict__ w, int c_out,
                                                    int c_in, int rows_p, int cols_p) {
  constexpr int G = Alg<DIM>::G;
  for (uint32_t i = threadIdx.x; i < G * 2 * (img_bytes >> 4); i += blockDim.x)
    reinterpret_cast<float4*>(img0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int total = c_out * c_in * G;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int g = idx % G, i = (idx / G) % c_in, o = idx / (G * c_in);
    const float x = w[idx];
    const float hi = tf32_hi(x);
    const int r = TRANS ? i : o, c = TRANS ? o : i;
    if (r < rows_p && c < cols_p) {
      const uint32_t off = plane_off(rows_p, r, c);
      *reinterpret_cast<float*>(img0 + (size_t)(g * 2 + 0) * img_bytes + off) = hi;
      *reinterpret_cast<float*>(img0 + (size_t)(g * 2 + 1) * img_bytes + off) = x - hi;
    }
  }
}

// A-operand descriptors of blade b in a chunk buffer half
__device__ __forceinline__ uint64_t chunk_desc(uint32_t half_saddr, int b) {
  return smem_desc(half_saddr + (uint32_t)b * kPS, kKH, 128u);
}

}  // namespace tcb
}  // namespace csmpn
