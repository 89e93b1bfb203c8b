// Row-local multivector kernels: geometric product, per-grade forms, MVSiLU, NormalizationLayer,
// MVLayerNorm, weighted geometric product (forward + backward), with deterministic parameter-gradient
// reductions (per-thread -> per-CTA (shared memory, fixed order) -> workspace -> fixed-order final sum).
//
// Layout: [rows, C, B] fp32, blades innermost.  One thread owns one (row, channel) multivector in registers
// (B <= 32 floats); a CTA is (C x RY) threads so the channel of a thread is fixed over its grid-stride loop.
// These kernels are HBM-bound (a few FLOP per byte): 128-bit loads/stores, no shared-memory staging of data.
#include "common.cuh"

namespace csmpn {

constexpr int kMaxPartialCtas = 1024;  // upper bound on CTAs that write parameter-gradient partials

// ---------------------------------------------------------------------------------------------------
// geometric product
template <int DIM, bool MET>
__global__ void __launch_bounds__(256) gp_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ out, int64_t n, int a_bcast, int b_bcast,
                                                     MetricParams mp) {
  using A = Alg<DIM>;
  constexpr int B = A::B;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float av[B], bv[B], ov[B];
    load_vec<B>(av, a + (a_bcast ? 0 : e * B));
    load_vec<B>(bv, b + (b_bcast ? 0 : e * B));
    A::template gp<MET>(av, bv, mp.mf, ov);
    store_vec<B>(out + e * B, ov);
  }
}

template <int DIM, bool MET>
__global__ void __launch_bounds__(256) gp_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     const float* __restrict__ go, float* __restrict__ ga,
                                                     float* __restrict__ gb, int64_t n, int a_bcast, int b_bcast,
                                                     MetricParams mp) {
  using A = Alg<DIM>;
  constexpr int B = A::B;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float av[B], bv[B], gv[B], gav[B], gbv[B];
    load_vec<B>(av, a + (a_bcast ? 0 : e * B));
    load_vec<B>(bv, b + (b_bcast ? 0 : e * B));
    load_vec<B>(gv, go + e * B);
    A::template gp_bwd<MET>(av, bv, gv, mp.mf, gav, gbv);
    store_vec<B>(ga + e * B, gav);
    store_vec<B>(gb + e * B, gbv);
  }
}

// ---------------------------------------------------------------------------------------------------
// per-grade forms: mode 0 -> q_g, mode 1 -> (q_g^2 + 1e-16)^(1/4)
template <int DIM>
__global__ void __launch_bounds__(256) forms_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n,
                                                        int mode, MetricParams mp) {
  using A = Alg<DIM>;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float xv[A::B], q[A::G];
    load_vec<A::B>(xv, x + e * A::B);
    grade_q<DIM>(xv, mp.qs, q);
#pragma unroll
    for (int g = 0; g < A::G; ++g) out[e * A::G + g] = mode ? smooth_abs_sqrt(q[g]) : q[g];
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) forms_bwd_kernel(const float* __restrict__ x, const float* __restrict__ go,
                                                        float* __restrict__ gx, int64_t n, int mode, MetricParams mp) {
  using A = Alg<DIM>;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float xv[A::B], q[A::G], dq[A::G], gv[A::B];
    load_vec<A::B>(xv, x + e * A::B);
    grade_q<DIM>(xv, mp.qs, q);
#pragma unroll
    for (int g = 0; g < A::G; ++g) {
      float d = go[e * A::G + g];
      if (mode) {  // d norm / d q = q / (2 norm^3)
        float nn = smooth_abs_sqrt(q[g]);
        d = d * q[g] / (2.f * nn * nn * nn);
      }
      dq[g] = d;
    }
#pragma unroll
    for (int i = 0; i < A::B; ++i) gv[i] = 2.f * mp.qs[i] * xv[i] * dq[A::grade_of(i)];
    store_vec<A::B>(gx + e * A::B, gv);
  }
}

// ---------------------------------------------------------------------------------------------------
// helper: deterministic CTA reduction of per-thread parameter-gradient accumulators.
// threads are (n = tid % C, ry = tid / C); acc[NP] per thread belongs to channel n.
// partial layout in the workspace: ws[cta][n * NP + p]
template <int NP>
__device__ __forceinline__ void cta_param_reduce(const float (&acc)[NP], float* __restrict__ ws_cta, int C, int RY,
                                                 float* smem /* blockDim floats */) {
  const int tid = threadIdx.x, n = tid % C, ry = tid / C;
  const bool live = ry < RY;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    __syncthreads();
    if (live) smem[ry * C + n] = acc[p];
    __syncthreads();
    if (ry == 0 && n < C) {
      float s = 0.f;
      for (int y = 0; y < RY; ++y) s += smem[y * C + n];
      ws_cta[n * NP + p] = s;
    }
  }
}

// out[q] = sum_cta ws[cta][q]   (fixed order)
__global__ void partial_sum_kernel(const float* __restrict__ ws, float* __restrict__ out, int n_params, int n_ctas) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_params) return;
  float s = 0.f;
  for (int c = 0; c < n_ctas; ++c) s += ws[(size_t)c * n_params + q];
  out[q] = s;
}

struct RowGrid {
  int C, RY, threads, grid;
  int64_t iters;  // uniform iteration count of the row loop
};

inline RowGrid make_row_grid(int64_t rows, int C, int max_ctas) {
  RowGrid g;
  g.C = C;
  g.RY = 256 / C > 0 ? 256 / C : 1;
  g.threads = ((C * g.RY + 31) / 32) * 32;
  int64_t tiles = (rows + g.RY - 1) / g.RY;
  int64_t want = tiles < max_ctas ? tiles : max_ctas;
  g.grid = (int)(want > 0 ? want : 1);
  g.iters = (tiles + g.grid - 1) / g.grid;
  return g;
}

// ---------------------------------------------------------------------------------------------------
// MVSiLU
template <int DIM>
__global__ void __launch_bounds__(256) mvsilu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                          const float* __restrict__ pb, float* __restrict__ y,
                                                          int64_t rows, int C, int RY, int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  if (ry >= RY) return;
  float a[A::G], b[A::G];
#pragma unroll
  for (int g = 0; g < A::G; ++g) { a[g] = pa[n * A::G + g]; b[g] = pb[n * A::G + g]; }
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (r >= rows) continue;
    float xv[A::B], q[A::G], s[A::G];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    grade_q<DIM>(xv, mp.qs, q);
    q[0] = xv[0];
#pragma unroll
    for (int g = 0; g < A::G; ++g) s[g] = sigmoidf_(fmaf(a[g], q[g], b[g]));
#pragma unroll
    for (int i = 0; i < A::B; ++i) xv[i] *= s[A::grade_of(i)];
    store_vec<A::B>(y + off, xv);
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) mvsilu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                          const float* __restrict__ pb, const float* __restrict__ gy,
                                                          float* __restrict__ gx, float* __restrict__ ws, int64_t rows,
                                                          int C, int RY, int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  extern __shared__ float smem[];
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  const bool live = ry < RY;
  float a[A::G], b[A::G], acc[2 * A::G];
#pragma unroll
  for (int g = 0; g < A::G; ++g) {
    a[g] = live ? pa[n * A::G + g] : 0.f;
    b[g] = live ? pb[n * A::G + g] : 0.f;
    acc[g] = 0.f; acc[A::G + g] = 0.f;
  }
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (!live || r >= rows) continue;
    float xv[A::B], gv[A::B], q[A::G], t[A::G];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    load_vec<A::B>(gv, gy + off);
    grade_q<DIM>(xv, mp.qs, q);
    q[0] = xv[0];
#pragma unroll
    for (int g = 0; g < A::G; ++g) t[g] = 0.f;
#pragma unroll
    for (int i = 0; i < A::B; ++i) t[A::grade_of(i)] = fmaf(gv[i], xv[i], t[A::grade_of(i)]);
    float sg[A::G], ds[A::G];
#pragma unroll
    for (int g = 0; g < A::G; ++g) {
      sg[g] = sigmoidf_(fmaf(a[g], q[g], b[g]));
      ds[g] = t[g] * sg[g] * (1.f - sg[g]);  // d loss / d pre-activation
      acc[g] = fmaf(ds[g], q[g], acc[g]);
      acc[A::G + g] += ds[g];
    }
#pragma unroll
    for (int i = 0; i < A::B; ++i) {
      const int g = A::grade_of(i);
      float dinv = (g == 0) ? 1.f : 2.f * mp.qs[i] * xv[i];
      gv[i] = fmaf(sg[g], gv[i], ds[g] * a[g] * dinv);
    }
    store_vec<A::B>(gx + off, gv);
  }
  cta_param_reduce<2 * A::G>(acc, ws + (size_t)blockIdx.x * C * 2 * A::G, C, RY, smem);
}

// scatter [C][2G] partial layout into grad_a[C][G], grad_b[C][G]
__global__ void silu_grad_final_kernel(const float* __restrict__ ws, float* __restrict__ ga, float* __restrict__ gb,
                                       int C, int G, int n_ctas) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= C * 2 * G) return;
  float s = 0.f;
  for (int c = 0; c < n_ctas; ++c) s += ws[(size_t)c * C * 2 * G + q];
  int n = q / (2 * G), p = q % (2 * G);
  if (p < G) ga[n * G + p] = s; else gb[n * G + p - G] = s;
}

// ---------------------------------------------------------------------------------------------------
// NormalizationLayer
template <int DIM>
__global__ void __launch_bounds__(256) mvnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                          float* __restrict__ y, int64_t rows, int C, int RY,
                                                          int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  if (ry >= RY) return;
  float sa[A::G];
#pragma unroll
  for (int g = 0; g < A::G; ++g) sa[g] = sigmoidf_(pa[n * A::G + g]);
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (r >= rows) continue;
    float xv[A::B], q[A::G], inv[A::G];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    grade_q<DIM>(xv, mp.qs, q);
#pragma unroll
    for (int g = 0; g < A::G; ++g) inv[g] = 1.f / (fmaf(sa[g], smooth_abs_sqrt(q[g]) - 1.f, 1.f) + kEps);
#pragma unroll
    for (int i = 0; i < A::B; ++i) xv[i] *= inv[A::grade_of(i)];
    store_vec<A::B>(y + off, xv);
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) mvnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                          const float* __restrict__ gy, float* __restrict__ gx,
                                                          float* __restrict__ ws, int64_t rows, int C, int RY,
                                                          int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  extern __shared__ float smem[];
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  const bool live = ry < RY;
  float sa[A::G], acc[A::G];
#pragma unroll
  for (int g = 0; g < A::G; ++g) { sa[g] = live ? sigmoidf_(pa[n * A::G + g]) : 0.f; acc[g] = 0.f; }
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (!live || r >= rows) continue;
    float xv[A::B], gv[A::B], q[A::G], t[A::G], inv[A::G], coef[A::G];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    load_vec<A::B>(gv, gy + off);
    grade_q<DIM>(xv, mp.qs, q);
#pragma unroll
    for (int g = 0; g < A::G; ++g) t[g] = 0.f;
#pragma unroll
    for (int i = 0; i < A::B; ++i) t[A::grade_of(i)] = fmaf(gv[i], xv[i], t[A::grade_of(i)]);
#pragma unroll
    for (int g = 0; g < A::G; ++g) {
      float nn = smooth_abs_sqrt(q[g]);
      inv[g] = 1.f / (fmaf(sa[g], nn - 1.f, 1.f) + kEps);
      float dd = -t[g] * inv[g] * inv[g];               // d loss / d denominator
      acc[g] = fmaf(dd * (nn - 1.f), sa[g] * (1.f - sa[g]), acc[g]);
      coef[g] = dd * sa[g] * q[g] / (nn * nn * nn);     // times qs_i x_i = d denom / d x_i
    }
#pragma unroll
    for (int i = 0; i < A::B; ++i) {
      const int g = A::grade_of(i);
      gv[i] = fmaf(gv[i], inv[g], coef[g] * mp.qs[i] * xv[i]);
    }
    store_vec<A::B>(gx + off, gv);
  }
  cta_param_reduce<A::G>(acc, ws + (size_t)blockIdx.x * C * A::G, C, RY, smem);
}

// ---------------------------------------------------------------------------------------------------
// MVLayerNorm: one CTA iteration handles RY rows x C channels; the channel mean goes through shared memory.
template <int DIM>
__device__ __forceinline__ float full_q(const float* xv, const float* qs) {
  float Q = 0.f;
#pragma unroll
  for (int i = 0; i < Alg<DIM>::B; ++i) Q = fmaf(qs[i] * xv[i], xv[i], Q);
  return Q;
}

template <int DIM>
__global__ void __launch_bounds__(256) mvln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                        float* __restrict__ y, int64_t rows, int C, int RY,
                                                        int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  extern __shared__ float smem[];  // [RY*C] norms + [RY] mu
  float* mu = smem + RY * C;
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  const bool live = ry < RY;
  const float a = live ? pa[n] : 0.f;
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    const bool ok = live && r < rows;
    float xv[A::B];
    const int64_t off = (r * C + n) * A::B;
    if (ok) {
      load_vec<A::B>(xv, x + off);
      smem[ry * C + n] = smooth_abs_sqrt(full_q<DIM>(xv, mp.qs));
    }
    __syncthreads();
    if (ok && n == 0) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += smem[ry * C + c];
      mu[ry] = s / (float)C + kEps;
    }
    __syncthreads();
    if (ok) {
      const float sc = a / mu[ry];
#pragma unroll
      for (int i = 0; i < A::B; ++i) xv[i] *= sc;
      store_vec<A::B>(y + off, xv);
    }
    __syncthreads();
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) mvln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pa,
                                                        const float* __restrict__ gy, float* __restrict__ gx,
                                                        float* __restrict__ ws, int64_t rows, int C, int RY,
                                                        int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  extern __shared__ float smem[];  // [RY*C] norms, [RY*C] dots, [RY] mu, [RY] dmu ; later reused by the reduce
  float* s_nu = smem;
  float* s_dot = smem + RY * C;
  float* mu = smem + 2 * RY * C;
  float* dmu = mu + RY;
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  const bool live = ry < RY;
  const float a = live ? pa[n] : 0.f;
  float acc[1] = {0.f};
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    const bool ok = live && r < rows;
    float xv[A::B], gv[A::B];
    float Q = 0.f, nu = 1.f, dot = 0.f;
    const int64_t off = (r * C + n) * A::B;
    if (ok) {
      load_vec<A::B>(xv, x + off);
      load_vec<A::B>(gv, gy + off);
      Q = full_q<DIM>(xv, mp.qs);
      nu = smooth_abs_sqrt(Q);
#pragma unroll
      for (int i = 0; i < A::B; ++i) dot = fmaf(gv[i], xv[i], dot);
      s_nu[ry * C + n] = nu;
      s_dot[ry * C + n] = a * dot;
    }
    __syncthreads();
    if (ok && n == 0) {
      float s = 0.f, d = 0.f;
      for (int c = 0; c < C; ++c) { s += s_nu[ry * C + c]; d += s_dot[ry * C + c]; }
      float m = s / (float)C + kEps;
      mu[ry] = m;
      dmu[ry] = -d / (m * m);
    }
    __syncthreads();
    if (ok) {
      const float m = mu[ry];
      acc[0] += dot / m;
      const float k1 = a / m;
      const float k2 = dmu[ry] / (float)C * Q / (nu * nu * nu);  // d mu / d x_i = (1/C) Q qs_i x_i / nu^3
#pragma unroll
      for (int i = 0; i < A::B; ++i) gv[i] = fmaf(k1, gv[i], k2 * mp.qs[i] * xv[i]);
      store_vec<A::B>(gx + off, gv);
    }
    __syncthreads();
  }
  cta_param_reduce<1>(acc, ws + (size_t)blockIdx.x * C, C, RY, smem);
}

// ---------------------------------------------------------------------------------------------------
// weighted geometric product
template <int DIM, bool MET>
__global__ void __launch_bounds__(256) wgp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ rr,
                                                       const float* __restrict__ w, const float* __restrict__ left,
                                                       float scale, float* __restrict__ out, int64_t rows, int C,
                                                       int RY, int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  if (ry >= RY) return;
  float wv[A::P];
#pragma unroll
  for (int p = 0; p < A::P; ++p) wv[p] = w[n * A::P + p];
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (r >= rows) continue;
    float xv[A::B], rv[A::B], z[A::B];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    load_vec<A::B>(rv, rr + off);
    if (left) load_vec<A::B>(z, left + off);
    else {
#pragma unroll
      for (int i = 0; i < A::B; ++i) z[i] = 0.f;
    }
    A::template wgp<MET>(xv, rv, wv, mp.mf, z);
#pragma unroll
    for (int i = 0; i < A::B; ++i) z[i] *= scale;
    store_vec<A::B>(out + off, z);
  }
}

template <int DIM, bool MET>
__global__ void __launch_bounds__(256) wgp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ rr,
                                                       const float* __restrict__ w, const float* __restrict__ go,
                                                       float scale, float* __restrict__ gx, float* __restrict__ gr,
                                                       float* __restrict__ ws, int64_t rows, int C, int RY,
                                                       int64_t iters, MetricParams mp) {
  using A = Alg<DIM>;
  extern __shared__ float smem[];
  const int n = threadIdx.x % C, ry = threadIdx.x / C;
  const bool live = ry < RY;
  float wv[A::P], dw[A::P];
#pragma unroll
  for (int p = 0; p < A::P; ++p) { wv[p] = live ? w[n * A::P + p] : 0.f; dw[p] = 0.f; }
  for (int64_t it = 0; it < iters; ++it) {
    int64_t r = (it * gridDim.x + blockIdx.x) * RY + ry;
    if (!live || r >= rows) continue;
    float xv[A::B], rv[A::B], dz[A::B], dx[A::B], dr[A::B];
    const int64_t off = (r * C + n) * A::B;
    load_vec<A::B>(xv, x + off);
    load_vec<A::B>(rv, rr + off);
    load_vec<A::B>(dz, go + off);
#pragma unroll
    for (int i = 0; i < A::B; ++i) { dz[i] *= scale; dx[i] = 0.f; dr[i] = 0.f; }
    A::template wgp_bwd<MET>(xv, rv, wv, dz, mp.mf, dx, dr, dw);
    store_vec<A::B>(gx + off, dx);
    store_vec<A::B>(gr + off, dr);
  }
  cta_param_reduce<A::P>(dw, ws + (size_t)blockIdx.x * C * A::P, C, RY, smem);
}

// ===================================================================================================
// host launchers
inline int ew_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count_cached() * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace csmpn

using namespace csmpn;

extern "C" {

int64_t csmpn_param_grad_workspace(int64_t n_params) {
  if (n_params < 0) return 0;
  return (int64_t)kMaxPartialCtas * n_params * (int64_t)sizeof(float);
}

int csmpn_gp_fwd(int dim, const float* metric, const float* a, const float* b, float* out, int64_t n_mv, int a_bcast,
                 int b_bcast, csmpn_stream_t stream) {
  MetricParams mp;
  int st = make_metric_params(dim, metric, &mp);
  if (st) return st;
  if (n_mv < 0 || (n_mv > 0 && (!a || !b || !out))) return CSMPN_ERR_BAD_ARG;
  if (n_mv == 0) return CSMPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_DISPATCH_DIM(dim, D, {
    if (mp.euclid) gp_fwd_kernel<D, false><<<ew_grid(n_mv), 256, 0, s>>>(a, b, out, n_mv, a_bcast, b_bcast, mp);
    else gp_fwd_kernel<D, true><<<ew_grid(n_mv), 256, 0, s>>>(a, b, out, n_mv, a_bcast, b_bcast, mp);
  });
  CSMPN_LAUNCH_CHECK("gp_fwd");
  return CSMPN_OK;
}

int csmpn_gp_bwd(int dim, const float* metric, const float* a, const float* b, const float* grad_out, float* grad_a,
                 float* grad_b, int64_t n_mv, int a_bcast, int b_bcast, csmpn_stream_t stream) {
  MetricParams mp;
  int st = make_metric_params(dim, metric, &mp);
  if (st) return st;
  if (n_mv < 0 || (n_mv > 0 && (!a || !b || !grad_out || !grad_a || !grad_b))) return CSMPN_ERR_BAD_ARG;
  if (n_mv == 0) return CSMPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_DISPATCH_DIM(dim, D, {
    if (mp.euclid)
      gp_bwd_kernel<D, false><<<ew_grid(n_mv), 256, 0, s>>>(a, b, grad_out, grad_a, grad_b, n_mv, a_bcast, b_bcast, mp);
    else
      gp_bwd_kernel<D, true><<<ew_grid(n_mv), 256, 0, s>>>(a, b, grad_out, grad_a, grad_b, n_mv, a_bcast, b_bcast, mp);
  });
  CSMPN_LAUNCH_CHECK("gp_bwd");
  return CSMPN_OK;
}

int csmpn_grade_forms_fwd(int dim, const float* metric, const float* x, float* out, int64_t n_mv, int mode,
                          csmpn_stream_t stream) {
  MetricParams mp;
  int st = make_metric_params(dim, metric, &mp);
  if (st) return st;
  if (n_mv < 0 || (n_mv > 0 && (!x || !out))) return CSMPN_ERR_BAD_ARG;
  if (n_mv == 0) return CSMPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_DISPATCH_DIM(dim, D, { forms_fwd_kernel<D><<<ew_grid(n_mv), 256, 0, s>>>(x, out, n_mv, mode, mp); });
  CSMPN_LAUNCH_CHECK("grade_forms_fwd");
  return CSMPN_OK;
}

int csmpn_grade_forms_bwd(int dim, const float* metric, const float* x, const float* grad_out, float* grad_x,
                          int64_t n_mv, int mode, csmpn_stream_t stream) {
  MetricParams mp;
  int st = make_metric_params(dim, metric, &mp);
  if (st) return st;
  if (n_mv < 0 || (n_mv > 0 && (!x || !grad_out || !grad_x))) return CSMPN_ERR_BAD_ARG;
  if (n_mv == 0) return CSMPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_DISPATCH_DIM(dim, D, { forms_bwd_kernel<D><<<ew_grid(n_mv), 256, 0, s>>>(x, grad_out, grad_x, n_mv, mode, mp); });
  CSMPN_LAUNCH_CHECK("grade_forms_bwd");
  return CSMPN_OK;
}

#define CSMPN_ROW_PROLOGUE(nparams_per_channel)                                              \
  MetricParams mp;                                                                           \
  int st = make_metric_params(dim, metric, &mp);                                             \
  if (st) return st;                                                                         \
  if (rows < 0 || channels <= 0) return CSMPN_ERR_BAD_ARG;                                     \
  if (channels > 256) return CSMPN_ERR_UNSUPPORTED;                \
  cudaStream_t s = (cudaStream_t)stream;                                                     \
  RowGrid rg = make_row_grid(rows, channels, sm_count_cached() * 4 < kMaxPartialCtas ? sm_count_cached() * 4 : kMaxPartialCtas); \
  (void)s; (void)rg;

int csmpn_mvsilu_fwd(int dim, const float* metric, const float* x, const float* a, const float* b, float* y,
                     int64_t rows, int channels, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (rows == 0) return CSMPN_OK;
  if (!x || !a || !b || !y) return CSMPN_ERR_BAD_ARG;
  CSMPN_DISPATCH_DIM(dim, D, {
    mvsilu_fwd_kernel<D><<<rg.grid, rg.threads, 0, s>>>(x, a, b, y, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvsilu_fwd");
  return CSMPN_OK;
}

int csmpn_mvsilu_bwd(int dim, const float* metric, const float* x, const float* a, const float* b, const float* grad_y,
                     float* grad_x, float* grad_a, float* grad_b, int64_t rows, int channels, void* workspace,
                     int64_t workspace_bytes, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (!a || !b || !grad_a || !grad_b || (rows > 0 && (!x || !grad_y || !grad_x))) return CSMPN_ERR_BAD_ARG;
  const int G = dim + 1, np = channels * 2 * G;
  if (!workspace || workspace_bytes < (int64_t)rg.grid * np * 4) return CSMPN_ERR_WORKSPACE;
  float* ws = (float*)workspace;
  CSMPN_DISPATCH_DIM(dim, D, {
    mvsilu_bwd_kernel<D><<<rg.grid, rg.threads, rg.threads * sizeof(float), s>>>(x, a, b, grad_y, grad_x, ws, rows,
                                                                                  channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvsilu_bwd");
  silu_grad_final_kernel<<<(np + 127) / 128, 128, 0, s>>>(ws, grad_a, grad_b, channels, G, rg.grid);
  CSMPN_LAUNCH_CHECK("mvsilu_bwd_final");
  return CSMPN_OK;
}

int csmpn_mvnorm_fwd(int dim, const float* metric, const float* x, const float* a, float* y, int64_t rows, int channels,
                     csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (rows == 0) return CSMPN_OK;
  if (!x || !a || !y) return CSMPN_ERR_BAD_ARG;
  CSMPN_DISPATCH_DIM(dim, D, {
    mvnorm_fwd_kernel<D><<<rg.grid, rg.threads, 0, s>>>(x, a, y, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvnorm_fwd");
  return CSMPN_OK;
}

int csmpn_mvnorm_bwd(int dim, const float* metric, const float* x, const float* a, const float* grad_y, float* grad_x,
                     float* grad_a, int64_t rows, int channels, void* workspace, int64_t workspace_bytes,
                     csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (!a || !grad_a || (rows > 0 && (!x || !grad_y || !grad_x))) return CSMPN_ERR_BAD_ARG;
  const int np = channels * (dim + 1);
  if (!workspace || workspace_bytes < (int64_t)rg.grid * np * 4) return CSMPN_ERR_WORKSPACE;
  float* ws = (float*)workspace;
  CSMPN_DISPATCH_DIM(dim, D, {
    mvnorm_bwd_kernel<D><<<rg.grid, rg.threads, rg.threads * sizeof(float), s>>>(x, a, grad_y, grad_x, ws, rows,
                                                                                  channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvnorm_bwd");
  partial_sum_kernel<<<(np + 127) / 128, 128, 0, s>>>(ws, grad_a, np, rg.grid);
  CSMPN_LAUNCH_CHECK("mvnorm_bwd_final");
  return CSMPN_OK;
}

int csmpn_mvlayernorm_fwd(int dim, const float* metric, const float* x, const float* a, float* y, int64_t rows,
                          int channels, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (rows == 0) return CSMPN_OK;
  if (!x || !a || !y) return CSMPN_ERR_BAD_ARG;
  size_t sm = (size_t)(rg.RY * channels + rg.RY) * sizeof(float);
  CSMPN_DISPATCH_DIM(dim, D, {
    mvln_fwd_kernel<D><<<rg.grid, rg.threads, sm, s>>>(x, a, y, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvlayernorm_fwd");
  return CSMPN_OK;
}

int csmpn_mvlayernorm_bwd(int dim, const float* metric, const float* x, const float* a, const float* grad_y,
                          float* grad_x, float* grad_a, int64_t rows, int channels, void* workspace,
                          int64_t workspace_bytes, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (!a || !grad_a || (rows > 0 && (!x || !grad_y || !grad_x))) return CSMPN_ERR_BAD_ARG;
  const int np = channels;
  if (!workspace || workspace_bytes < (int64_t)rg.grid * np * 4) return CSMPN_ERR_WORKSPACE;
  float* ws = (float*)workspace;
  size_t sm = (size_t)(2 * rg.RY * channels + 2 * rg.RY) * sizeof(float);
  if (sm < rg.threads * sizeof(float)) sm = rg.threads * sizeof(float);
  CSMPN_DISPATCH_DIM(dim, D, {
    mvln_bwd_kernel<D><<<rg.grid, rg.threads, sm, s>>>(x, a, grad_y, grad_x, ws, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("mvlayernorm_bwd");
  partial_sum_kernel<<<(np + 127) / 128, 128, 0, s>>>(ws, grad_a, np, rg.grid);
  CSMPN_LAUNCH_CHECK("mvlayernorm_bwd_final");
  return CSMPN_OK;
}

int csmpn_wgp_fwd(int dim, const float* metric, const float* x, const float* r, const float* w, const float* left,
                  float scale, float* out, int64_t rows, int channels, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (rows == 0) return CSMPN_OK;
  if (!x || !r || !w || !out) return CSMPN_ERR_BAD_ARG;
  CSMPN_DISPATCH_DIM(dim, D, {
    if (mp.euclid)
      wgp_fwd_kernel<D, false><<<rg.grid, rg.threads, 0, s>>>(x, r, w, left, scale, out, rows, channels, rg.RY, rg.iters, mp);
    else
      wgp_fwd_kernel<D, true><<<rg.grid, rg.threads, 0, s>>>(x, r, w, left, scale, out, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("wgp_fwd");
  return CSMPN_OK;
}

int csmpn_wgp_bwd(int dim, const float* metric, const float* x, const float* r, const float* w, const float* grad_out,
                  float scale, float* grad_x, float* grad_r, float* grad_w, int64_t rows, int channels, void* workspace,
                  int64_t workspace_bytes, csmpn_stream_t stream) {
  CSMPN_ROW_PROLOGUE(0)
  if (!w || !grad_w || (rows > 0 && (!x || !r || !grad_out || !grad_x || !grad_r))) return CSMPN_ERR_BAD_ARG;
  int P = 0;
  CSMPN_DISPATCH_DIM(dim, D, { P = Alg<D>::P; });
  const int np = channels * P;
  if (!workspace || workspace_bytes < (int64_t)rg.grid * np * 4) return CSMPN_ERR_WORKSPACE;
  float* ws = (float*)workspace;
  CSMPN_DISPATCH_DIM(dim, D, {
    if (mp.euclid)
      wgp_bwd_kernel<D, false><<<rg.grid, rg.threads, rg.threads * sizeof(float), s>>>(
          x, r, w, grad_out, scale, grad_x, grad_r, ws, rows, channels, rg.RY, rg.iters, mp);
    else
      wgp_bwd_kernel<D, true><<<rg.grid, rg.threads, rg.threads * sizeof(float), s>>>(
          x, r, w, grad_out, scale, grad_x, grad_r, ws, rows, channels, rg.RY, rg.iters, mp);
  });
  CSMPN_LAUNCH_CHECK("wgp_bwd");
  partial_sum_kernel<<<(np + 127) / 128, 128, 0, s>>>(ws, grad_w, np, rg.grid);
  CSMPN_LAUNCH_CHECK("wgp_bwd_final");
  return CSMPN_OK;
}

}  // extern "C"
