// Per-grade channel GEMM core shared by the MVLinear kernels and the fused CEMLP-block kernels.
//
//   y[r, o, i] = sum_k x[r, k, i] * W(k, o)[grade(i)]          (cegnn_utils.py:329-331, without materialising
//                                                                the repeat_interleave'd [Cout,Cin,B] weight)
//
// FP32 FMA formulation (the fp32 parity budget of 1e-5 rules out single-pass TF32/BF16 tensor-core math):
// a thread owns an RB x NCH x B register tile (RB rows, NCH output channels, all B blades of each).  Per k it
// issues RB*B/4 128-bit loads of x and NCH (GP/4) 128-bit loads of w against RB*NCH*B FFMAs.  Output channels
// are assigned with stride NC (= #channel groups) so that the 8..16 lanes of a warp that differ in channel
// group read consecutive weight rows (row stride = 4 mod 32 words -> conflict-free), and lanes that differ in
// row group read x rows whose stride is also 4 mod 32 words.
#pragma once
#include "common.cuh"

namespace csmpn {

template <int DIM> struct GemmCfg;
template <> struct GemmCfg<1> { static constexpr int RB = 4, NCH = 4, GP = 2; };
template <> struct GemmCfg<2> { static constexpr int RB = 4, NCH = 4, GP = 4; };
template <> struct GemmCfg<3> { static constexpr int RB = 2, NCH = 4, GP = 4; };
template <> struct GemmCfg<4> { static constexpr int RB = 1, NCH = 4, GP = 8; };
template <> struct GemmCfg<5> { static constexpr int RB = 1, NCH = 2, GP = 8; };

// smallest stride >= n with stride % 32 == 4 (words): consecutive rows start 4 banks apart
__host__ __device__ inline int pad_stride(int n) { return n + ((4 - (n % 32)) + 32) % 32; }

// acc[j][a][i] += sum_{k < K} xs[j*xstride + k*B + i] * wp[a][k*wk_stride + grade(i)]
template <int DIM, int NCH = GemmCfg<DIM>::NCH, int RB = GemmCfg<DIM>::RB>
__device__ __forceinline__ void gemm_accumulate(float (&acc)[RB][NCH][Alg<DIM>::B],
                                                const float* __restrict__ xs, int xstride,
                                                const float* const (&wp)[NCH], int wk_stride, int K) {
  using A = Alg<DIM>;
  using Cfg = GemmCfg<DIM>;
  constexpr int B = A::B, GP = Cfg::GP;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float xv[RB][B], wv[NCH][GP];
#pragma unroll
    for (int j = 0; j < RB; ++j) load_vec<B>(xv[j], xs + j * xstride + k * B);
#pragma unroll
    for (int a = 0; a < NCH; ++a) load_vec<GP>(wv[a], wp[a] + k * wk_stride);
#pragma unroll
    for (int j = 0; j < RB; ++j)
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] = fmaf(xv[j][i], wv[a][A::grade_of(i)], acc[j][a][i]);
  }
}

// Cooperative copy of a [rows_in_tile x kc channels] slab of x (global, row stride kdim*B) into shared memory
// (row stride sx).  Rows >= rows_valid are zero-filled.
template <int DIM>
__device__ __forceinline__ void stage_rows(float* __restrict__ xs, int sx, const float* __restrict__ x, int64_t row0,
                                           int64_t rows_total, int tr, int kdim, int k0, int kc) {
  constexpr int B = Alg<DIM>::B;
  constexpr int V = (B % 4 == 0) ? 4 : 2;
  const int vpr = kc * B / V;  // vectors per row
  for (int idx = threadIdx.x; idx < tr * vpr; idx += blockDim.x) {
    const int r = idx / vpr, v = idx - r * vpr;
    const int64_t gr = row0 + r;
    if constexpr (V == 4) {
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows_total) val = *reinterpret_cast<const float4*>(x + (gr * kdim + k0) * B + 4 * v);
      *reinterpret_cast<float4*>(xs + r * sx + 4 * v) = val;
    } else {
      float2 val = make_float2(0.f, 0.f);
      if (gr < rows_total) val = *reinterpret_cast<const float2*>(x + (gr * kdim + k0) * B + 2 * v);
      *reinterpret_cast<float2*>(xs + r * sx + 2 * v) = val;
    }
  }
}

// Weight staging.  Global weight is W[n][m][g] (g < gw; gw = G for subspaces=True, 1 otherwise -- then the
// single value is replicated over all grades).
//  TRANS = false (y = x W^T, k = m, o = n):  ws[n * sw + (m - k0) * GP + g],  n < n_out, m in [k0, k0+kc)
//  TRANS = true  (dx = dy W,  k = n, o = m):  ws[(n - k0) * sw + m * GP + g],  n in [k0, k0+kc), m < n_out
template <int DIM, bool TRANS>
__device__ __forceinline__ void stage_weights(float* __restrict__ ws, int sw, const float* __restrict__ w, int c_out,
                                              int c_in, int gw, int k0, int kc) {
  constexpr int G = Alg<DIM>::G, GP = GemmCfg<DIM>::GP;
  if constexpr (G == 4 && GP == 4) {
    if (gw == G) {  // one 128-bit copy per (n, m): coalesced along m
      const int rows = TRANS ? kc : c_out, cols = TRANS ? c_in : kc;
      for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
        const int n = idx / cols, m = idx - n * cols;
        const int gn = TRANS ? k0 + n : n, gm = TRANS ? m : k0 + m;
        *reinterpret_cast<float4*>(ws + n * sw + m * 4) = *reinterpret_cast<const float4*>(w + ((int64_t)gn * c_in + gm) * 4);
      }
      return;
    }
  }
  if constexpr (!TRANS) {
    const int per_n = kc * G;
    for (int idx = threadIdx.x; idx < c_out * per_n; idx += blockDim.x) {
      const int n = idx / per_n, rem = idx - n * per_n;
      const int m = rem / G, g = rem - m * G;
      ws[n * sw + m * GP + g] = w[((int64_t)n * c_in + (k0 + m)) * gw + (gw == 1 ? 0 : g)];
    }
  } else {
    const int per_n = c_in * G;
    for (int idx = threadIdx.x; idx < kc * per_n; idx += blockDim.x) {
      const int n = idx / per_n, rem = idx - n * per_n;
      const int m = rem / G, g = rem - m * G;
      ws[n * sw + m * GP + g] = w[((int64_t)(k0 + n) * c_in + m) * gw + (gw == 1 ? 0 : g)];
    }
  }
}

}  // namespace csmpn
