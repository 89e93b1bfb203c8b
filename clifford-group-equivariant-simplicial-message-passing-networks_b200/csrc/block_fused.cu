// Fused CEMLP-block kernels (forward and backward) for Euclidean Cl(2,0), Cl(3,0), Cl(5,0).
//
// One block = MVLinear -> MVSiLU -> SteerableGeometricProductLayer -> MVLayerNorm  (cegnn_utils.py:180-207).
//
// Execution model (B200): persistent CTAs, one per SM, each looping over tiles of TR rows.
//  * The three weight matrices of the block stay RESIDENT in shared memory for the life of the CTA (natural
//    [n][m][g] layout, row stride = 4 mod 32 words, which serves both the forward and the transposed GEMMs
//    conflict-free).  When they do not fit (wide layers) they are staged per GEMM in K-chunks instead.
//  * The input rows of a tile are fetched by TMA: one cp.async.bulk per (row, source) into shared memory,
//    completion on an mbarrier.  For the EGCL message block the receiver rows h[dst] land in bufA and the sender
//    rows h[src] in bufB, and one shared-memory pass forms h[dst] - h[src]; the [E, C+2T, B] message input of the
//    reference (cegnn_utils.py:254-259) is never materialised in HBM.
//  * The three per-grade channel GEMMs read their K operand from shared memory and keep their outputs in registers,
//    where bias, MVSiLU, normalisation, the table-driven weighted geometric product and the layer norm are applied.
//
// Backward uses three saved [rows, C, B] tensors (pre-SiLU y1, pre-normalisation xr, pre-LayerNorm o) instead of
// recomputing the forward GEMMs: the layer is bound by the FP32 pipe, not by HBM, so three extra tensors of traffic
// are cheaper than a third of the backward FLOPs.  Weight gradients are tile-local GEMMs accumulated into per-CTA
// global accumulators; the small per-channel parameter gradients are reduced with fixed-order warp shuffles.  A last
// kernel sums the per-CTA partials in a fixed order, so every gradient is bit-reproducible run to run.
#include "gemm.cuh"

namespace csmpn {

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr int kMaxSmem = 220 * 1024;

struct FusedPlan {
  int tr, rg, nc, threads, nwarps;
  int sa, sb;          // row strides (words) of bufA (max(c_in, c) channels) and bufB
  int resident;        // weights resident in shared memory
  int sw1, swc;        // resident row strides: pad(c_in*GP), pad(c*GP)
  int off_w1, off_wr, off_wl;  // word offsets of the resident weights inside the weight area
  int kc1, kcc, kt1, ktc;      // staged mode: K-chunks (forward GEMM1 / CxC, transposed W1 / CxC)
  int wbuf;            // words of the weight area
  size_t smem;
  int grid;
  int ks;              // max row splits of the weight-gradient GEMMs
};

template <int DIM>
inline int make_fused_plan(const csmpn_block_desc& d, FusedPlan* out) {
  using Cfg = GemmCfg<DIM>;
  constexpr int B = Alg<DIM>::B, GP = Cfg::GP;
  FusedPlan p;
  memset(&p, 0, sizeof(p));
  const int c = d.c, cin = d.c0 + d.c1 + d.c2;
  p.nc = (c + Cfg::NCH - 1) / Cfg::NCH;
  if (p.nc > 256) return CSMPN_ERR_UNSUPPORTED;
  const int wide = cin > c ? cin : c;
  const int bwide = (d.mode == 1 && d.c0 > c) ? d.c0 : c;
  p.sa = pad_stride(wide * B);
  p.sb = pad_stride(bwide * B);
  p.sw1 = pad_stride(cin * GP);
  p.swc = pad_stride(c * GP);
  const int res_words = c * p.sw1 + 2 * c * p.swc;
  auto tile_words = [&](int tr) { return (size_t)tr * (p.sa + p.sb) + 2 * (size_t)tr * p.nc + 16; };
  // largest tile (<= 256 threads) that fits next to the resident weights
  int best = 0;
  for (int tr = 64; tr >= Cfg::RB; tr >>= 1) {
    if (tr % Cfg::RB) continue;
    const int rg = tr / Cfg::RB;
    if (((rg * p.nc + 31) / 32) * 32 > 256) continue;
    if ((tile_words(tr) + res_words) * 4 <= (size_t)kMaxSmem) { best = tr; break; }
  }
  if (best) {
    p.resident = 1;
    p.tr = best;
    p.off_w1 = 0;
    p.off_wr = c * p.sw1;
    p.off_wl = p.off_wr + c * p.swc;
    p.wbuf = res_words;
  } else {
    // staged weights: 22 KB staging area, ~128-thread tiles
    p.resident = 0;
    const int wb_words = 5632;
    int rg = 128 / p.nc;
    if (rg < 1) rg = 1;
    int max_rg = (DIM <= 3 ? 32 : 8) / Cfg::RB;
    if (rg > max_rg) rg = max_rg;
    p.tr = rg * Cfg::RB;
    auto fit_fwd = [&](int odim, int kdim) {
      int kc = (wb_words / odim - 4) / GP;
      if (kc > kdim) kc = kdim;
      return kc < 1 ? 1 : kc;
    };
    auto fit_trans = [&](int odim, int kdim) {
      int kc = wb_words / pad_stride(odim * GP);
      if (kc > kdim) kc = kdim;
      return kc < 1 ? 1 : kc;
    };
    p.kc1 = fit_fwd(c, cin);
    p.kcc = fit_fwd(c, c);
    p.kt1 = fit_trans(cin, c);
    p.ktc = fit_trans(c, c);
    int w1 = c * pad_stride(p.kc1 * GP), w2 = c * pad_stride(p.kcc * GP);
    int w3 = p.kt1 * pad_stride(cin * GP), w4 = p.ktc * pad_stride(c * GP);
    p.wbuf = w1 > w2 ? w1 : w2;
    if (w3 > p.wbuf) p.wbuf = w3;
    if (w4 > p.wbuf) p.wbuf = w4;
  }
  p.rg = p.tr / Cfg::RB;
  p.threads = ((p.rg * p.nc + 31) / 32) * 32;
  if (p.threads > 256) return CSMPN_ERR_UNSUPPORTED;
  p.nwarps = p.threads / 32;
  p.smem = (tile_words(p.tr) + p.wbuf) * sizeof(float);
  if (p.smem > (size_t)kMaxSmem) return CSMPN_ERR_UNSUPPORTED;
  int64_t tiles = (d.rows + p.tr - 1) / p.tr;
  int per_sm = (int)((size_t)(226 * 1024) / (p.smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  if (per_sm * p.threads > 512) per_sm = 512 / p.threads > 0 ? 512 / p.threads : 1;
  int64_t cap = (int64_t)sm_count_cached() * per_sm;
  p.grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
  // weight-gradient GEMMs: thread tiles of NA x 4 channels; rows of a tile split ks ways
  const int NA = Alg<DIM>::G <= 4 ? 4 : 2;
  int tc = ((c + NA - 1) / NA) * ((c + 3) / 4);
  p.ks = p.threads / tc > 0 ? p.threads / tc : 1;
  if (p.ks > 16) p.ks = 16;
  while (p.ks > 1 && p.tr % p.ks) --p.ks;
  *out = p;
  return CSMPN_OK;
}

// ---------------------------------------------------------------------------------------------------
// mbarrier / TMA (bulk async copy) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_row_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// thread-local multivector math (Euclidean: q_g = sum of squares)
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

// ---------------------------------------------------------------------------------------------------
// Tile input: TMA row copies (issued by warp 0) + one shared-memory pass.
//   mode 0:  bufA[r] = [ p0[row] | p1[row] | p2[row] ]
//   mode 1:  bufA[r] = [ p0[dst[row]] - p0[src[row]] | p1[eid[row]] ]     (sender rows pass through bufB)
// Caller guarantees every thread has executed fence_proxy_async() + __syncthreads() since the last generic access to
// bufA / bufB.  On return bufA is complete for the calling thread's own writes; a barrier must follow before other
// threads read it.
template <int DIM>
__device__ __forceinline__ void load_input_tile(float* __restrict__ bufA, int sa, float* __restrict__ bufB, int sb,
                                                const csmpn_block_desc& d, int64_t row0, int tr, uint64_t* bar,
                                                uint32_t parity) {
  constexpr int B = Alg<DIM>::B;
  const uint32_t b0 = d.c0 * B * 4, b1 = d.c1 * B * 4, b2 = d.c2 * B * 4;
  if (threadIdx.x < 32) {
    uint32_t bytes = 0;
    for (int r = threadIdx.x; r < tr; r += 32)
      if (row0 + r < d.rows) bytes += (d.mode == 1 ? 2 * b0 : b0) + b1 + b2;
    mbar_arrive_expect_tx(bar, bytes);
    for (int r = threadIdx.x; r < tr; r += 32) {
      const int64_t gr = row0 + r;
      if (gr >= d.rows) continue;
      float* ra = bufA + r * sa;
      if (d.mode == 0) {
        tma_row_g2s(ra, d.p0 + gr * (int64_t)d.c0 * B, b0, bar);
        if (b1) tma_row_g2s(ra + d.c0 * B, d.p1 + gr * (int64_t)d.c1 * B, b1, bar);
        if (b2) tma_row_g2s(ra + (d.c0 + d.c1) * B, d.p2 + gr * (int64_t)d.c2 * B, b2, bar);
      } else {
        const int64_t ri = __ldg(d.dst + gr), rj = __ldg(d.src + gr);
        tma_row_g2s(ra, d.p0 + ri * (int64_t)d.c0 * B, b0, bar);
        tma_row_g2s(bufB + r * sb, d.p0 + rj * (int64_t)d.c0 * B, b0, bar);
        if (b1) tma_row_g2s(ra + d.c0 * B, d.p1 + (int64_t)__ldg(d.eid + gr) * d.c1 * B, b1, bar);
      }
    }
  }
  mbar_wait(bar, parity);
  const int valid = (d.rows - row0) < tr ? (int)(d.rows - row0) : tr;
  if (d.mode == 1) {
    const int v0 = d.c0 * B / 4;
    for (int idx = threadIdx.x; idx < valid * v0; idx += blockDim.x) {
      const int r = idx / v0, v = idx - r * v0;
      float4 a = *reinterpret_cast<const float4*>(bufA + r * sa + 4 * v);
      const float4 b = *reinterpret_cast<const float4*>(bufB + r * sb + 4 * v);
      a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w;
      *reinterpret_cast<float4*>(bufA + r * sa + 4 * v) = a;
    }
  }
  if (valid < tr) {  // tail tile: zero the missing rows
    const int vpr = (d.c0 + d.c1 + d.c2) * B / 4;
    for (int idx = threadIdx.x; idx < (tr - valid) * vpr; idx += blockDim.x) {
      const int r = valid + idx / vpr, v = idx % vpr;
      *reinterpret_cast<float4*>(bufA + r * sa + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// weight operand of a GEMM stage: resident in shared memory, or staged from global in K-chunks
struct WRef {
  const float* g;   // global [c_out][c_in][G]
  float* s;         // shared: resident matrix, or the staging area
  int c_out, c_in;
  int sw;           // resident row stride
  int kchunk;       // staged mode
  int resident;
};

// acc += buf[:, 0:kdim] * W   for output channels  obase + c + a * nco  (a < NCHT), masked by owidth.
// The leading barrier publishes the shared-memory writes of the previous stage.
template <int DIM, bool TRANS, int NCHT>
__device__ __forceinline__ void gemm_run(float (&acc)[GemmCfg<DIM>::RB][NCHT][Alg<DIM>::B],
                                         const float* __restrict__ buf_rows, int stride, const WRef& w, bool active, int c,
                                         int nco, int obase, int owidth) {
  using Cfg = GemmCfg<DIM>;
  constexpr int B = Alg<DIM>::B, GP = Cfg::GP;
  const int kdim = TRANS ? w.c_out : w.c_in;
  const int odim = TRANS ? w.c_in : w.c_out;
  if (w.resident) {
    __syncthreads();
    if (active) {
      const float* wp[NCHT];
#pragma unroll
      for (int a = 0; a < NCHT; ++a) {
        int ol = c + a * nco;
        int o = obase + (ol < owidth ? ol : 0);
        wp[a] = TRANS ? w.s + o * GP : w.s + o * w.sw;
      }
      gemm_accumulate<DIM, NCHT>(acc, buf_rows, stride, wp, TRANS ? w.sw : GP, kdim);
    }
    return;
  }
  for (int k0 = 0; k0 < kdim; k0 += w.kchunk) {
    const int kc = (kdim - k0) < w.kchunk ? (kdim - k0) : w.kchunk;
    const int sw = TRANS ? pad_stride(odim * GP) : pad_stride(kc * GP);
    __syncthreads();
    stage_weights<DIM, TRANS>(w.s, sw, w.g, w.c_out, w.c_in, Alg<DIM>::G, k0, kc);
    __syncthreads();
    if (active) {
      const float* wp[NCHT];
#pragma unroll
      for (int a = 0; a < NCHT; ++a) {
        int ol = c + a * nco;
        int o = obase + (ol < owidth ? ol : 0);
        wp[a] = TRANS ? w.s + o * GP : w.s + o * sw;
      }
      gemm_accumulate<DIM, NCHT>(acc, buf_rows + k0 * B, stride, wp, TRANS ? sw : GP, kc);
    }
  }
}

template <int DIM>
__device__ __forceinline__ void setup_weights(const csmpn_block_desc& d, const FusedPlan& p, float* wsm, WRef& w1, WRef& wr,
                                              WRef& wl) {
  const int C = d.c, cin = d.c0 + d.c1 + d.c2;
  w1 = WRef{d.w1, p.resident ? wsm + p.off_w1 : wsm, C, cin, p.sw1, 0, p.resident};
  wr = WRef{d.wr, p.resident ? wsm + p.off_wr : wsm, C, C, p.swc, 0, p.resident};
  wl = WRef{d.wl, p.resident ? wsm + p.off_wl : wsm, C, C, p.swc, 0, p.resident};
  if (p.resident) {
    stage_weights<DIM, false>(w1.s, p.sw1, d.w1, C, cin, Alg<DIM>::G, 0, cin);
    stage_weights<DIM, false>(wr.s, p.swc, d.wr, C, C, Alg<DIM>::G, 0, C);
    stage_weights<DIM, false>(wl.s, p.swc, d.wl, C, C, Alg<DIM>::G, 0, C);
  }
}

// ===================================================================================================
// forward
template <int DIM>
__global__ void __launch_bounds__(256, 1) block_fwd_kernel(csmpn_block_desc d, FusedPlan p) {
  using A = Alg<DIM>;
  using Cfg = GemmCfg<DIM>;
  constexpr int B = A::B, G = A::G, P = A::P, RB = Cfg::RB, NCH = Cfg::NCH;
  extern __shared__ __align__(128) float smem[];
  float* bufA = smem;
  float* bufB = bufA + p.tr * p.sa;
  float* wsm = bufB + p.tr * p.sb;
  float* part = wsm + p.wbuf;  // [tr][nc] layer-norm partials
  uint64_t* bar = reinterpret_cast<uint64_t*>(part + 2 * p.tr * p.nc);
  const int C = d.c;
  const int c = threadIdx.x % p.nc, rg = threadIdx.x / p.nc;
  const bool active = rg < p.rg;
  int och[NCH];
  bool ov[NCH];
#pragma unroll
  for (int a = 0; a < NCH; ++a) { och[a] = c + a * p.nc; ov[a] = active && och[a] < C; if (!ov[a]) och[a] = 0; }
  const int64_t tiles = (d.rows + p.tr - 1) / p.tr;
  WRef w1, wr, wl;
  setup_weights<DIM>(d, p, wsm, w1, wr, wl);
  if (!p.resident) { w1.kchunk = p.kc1; wr.kchunk = p.kcc; wl.kchunk = p.kcc; }
  if (threadIdx.x == 0) mbar_init(bar, 32);
  uint32_t phase = 0;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * p.tr;
    const int lr0 = rg * RB;  // first local row of this thread
    fence_proxy_async();
    __syncthreads();  // previous tile's generic accesses to bufA / bufB / part are done (and the barrier is initialised)
    load_input_tile<DIM>(bufA, p.sa, bufB, p.sb, d, row0, p.tr, bar, phase);
    phase ^= 1;

    float acc[RB][NCH][B];
    // ---- GEMM1 + bias + MVSiLU  -> y2 in bufB
#pragma unroll
    for (int j = 0; j < RB; ++j)
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] = 0.f;
    gemm_run<DIM, false, NCH>(acc, bufA + lr0 * p.sa, p.sa, w1, active, c, p.nc, 0, C);
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
      const int n = och[a];
      float sa_[G], sb_[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { sa_[g] = d.sa[n * G + g]; sb_[g] = d.sb[n * G + g]; }
      const float b1 = d.has_b1 ? d.b1[n] : 0.f;
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        acc[j][a][0] += b1;
        const int64_t r = row0 + lr0 + j;
        if (d.save_y1 && r < d.rows) store_vec<B>(d.save_y1 + (r * C + n) * B, acc[j][a]);
        float sg[G], inv[G];
        silu_gates<DIM>(acc[j][a], sa_, sb_, sg, inv);
        float y2[B];
#pragma unroll
        for (int i = 0; i < B; ++i) y2[i] = acc[j][a][i] * sg[A::grade_of(i)];
        store_vec<B>(bufB + (lr0 + j) * p.sb + n * B, y2);
      }
    }
    // ---- GEMM-R -> xr (own slot of bufA)
#pragma unroll
    for (int j = 0; j < RB; ++j)
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] = 0.f;
    gemm_run<DIM, false, NCH>(acc, bufB + lr0 * p.sb, p.sb, wr, active, c, p.nc, 0, C);
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int64_t r = row0 + lr0 + j;
        if (d.save_xr && r < d.rows) store_vec<B>(d.save_xr + (r * C + och[a]) * B, acc[j][a]);
        store_vec<B>(bufA + (lr0 + j) * p.sa + och[a] * B, acc[j][a]);
      }
    }
    // ---- GEMM-L + bias, normalisation of xr, weighted geometric product, 1/sqrt2  -> o (registers)
#pragma unroll
    for (int j = 0; j < RB; ++j)
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] = 0.f;
    gemm_run<DIM, false, NCH>(acc, bufB + lr0 * p.sb, p.sb, wl, active, c, p.nc, 0, C);
    float nu_sum[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) nu_sum[j] = 0.f;
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
      const int n = och[a];
      float s[G], wv[P];
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] = sigmoidf_(d.na[n * G + g]);
#pragma unroll
      for (int q = 0; q < P; ++q) wv[q] = d.wp[n * P + q];
      const float bl = d.bl[n];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        float xr[B], y2[B], q[G], nrm[G], rinv[G];
        load_vec<B>(xr, bufA + (lr0 + j) * p.sa + n * B);
        load_vec<B>(y2, bufB + (lr0 + j) * p.sb + n * B);
        norm_factors<DIM>(xr, s, q, nrm, rinv);
#pragma unroll
        for (int i = 0; i < B; ++i) xr[i] *= rinv[A::grade_of(i)];
        acc[j][a][0] += bl;
        A::template wgp<false>(y2, xr, wv, nullptr, acc[j][a]);
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] *= kInvSqrt2;
        const int64_t r = row0 + lr0 + j;
        if (d.save_o && r < d.rows) store_vec<B>(d.save_o + (r * C + n) * B, acc[j][a]);
        nu_sum[j] += smooth_abs_sqrt(mv_sumsq<DIM>(acc[j][a]));
      }
    }
    // ---- MVLayerNorm: mean over channels of the norms (fixed-order sum of the per-thread partials)
    if (active) {
#pragma unroll
      for (int j = 0; j < RB; ++j) part[(lr0 + j) * p.nc + c] = nu_sum[j];
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        float sum = 0.f;
        for (int cc = 0; cc < p.nc; ++cc) sum += part[(lr0 + j) * p.nc + cc];
        const float inv_mu = 1.f / (sum / (float)C + kEps);
        const int64_t r = row0 + lr0 + j;
        if (r >= d.rows) continue;
#pragma unroll
        for (int a = 0; a < NCH; ++a) {
          if (!ov[a]) continue;
          const float sc = d.la[och[a]] * inv_mu;
          float y[B];
#pragma unroll
          for (int i = 0; i < B; ++i) y[i] = acc[j][a][i] * sc;
          if (d.res) {
            float rv[B];
            load_vec<B>(rv, d.res + (r * C + och[a]) * B);
#pragma unroll
            for (int i = 0; i < B; ++i) y[i] += rv[i];
          }
          store_vec<B>(d.y + (r * C + och[a]) * B, y);
        }
      }
    }
  }
}

// ===================================================================================================
// backward helpers

// Fixed-order reduction over the lanes of a warp that own the same channel group, then accumulation into this
// warp's private global accumulators.  vals[a][q]: contribution of this thread to parameter q of channel och[a].
// Threads are numbered t = rg * nc + c, so the lanes sharing c inside a warp are lane, lane + nc, lane + 2 nc, ...
template <int NCH, int Q>
__device__ __forceinline__ void warp_reduce_store(float (&vals)[NCH][Q], float* __restrict__ dest, const int (&och)[NCH],
                                                  const bool (&ov)[NCH], int nc) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < NCH; ++a)
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float v = vals[a][q];
      float s = v;
      for (int off = nc; off < 32; off += nc) {
        const float t = __shfl_down_sync(0xffffffffu, v, off);
        if (lane + off < 32) s += t;
      }
      vals[a][q] = s;
    }
  if (lane < nc) {  // chain heads: the first lane of the warp holding each channel group
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
#pragma unroll
      for (int q = 0; q < Q; ++q) dest[och[a] * Q + q] += vals[a][q];
    }
  }
}

// Tile-local weight-gradient GEMM over one column block of the m operand:
//   gacc[split][n][m_off + m][g] += sum_{r in split} sum_{i in g} nbuf[r][n][i] * mbuf[r][m][i],  m < cm_block
template <int DIM>
__device__ __forceinline__ void dw_tile(const float* __restrict__ nbuf, int nstride, int cn_total,
                                        const float* __restrict__ mbuf, int mstride, int cm_block, int m_off, int m_ld,
                                        int tr, int ks_max, float* __restrict__ gacc) {
  using A = Alg<DIM>;
  constexpr int B = A::B, G = A::G;
  constexpr int NA = (G <= 4) ? 4 : 2, MA = 4;
  const int ncn = (cn_total + NA - 1) / NA, ncm = (cm_block + MA - 1) / MA;
  const int tiles = ncn * ncm;
  int ks = (int)blockDim.x / tiles;
  if (ks < 1) ks = 1;
  if (ks > ks_max) ks = ks_max;
  while (ks > 1 && tr % ks) --ks;
  const int rows_per = tr / ks;
  for (int item = threadIdx.x; item < tiles * ks; item += blockDim.x) {
    const int split = item / tiles, tt = item - split * tiles;
    const int cm = tt % ncm, cn = tt / ncm;
    int nl[NA], ml[MA];
    bool nv[NA], mv[MA];
#pragma unroll
    for (int a = 0; a < NA; ++a) { int n = cn + a * ncn; nv[a] = n < cn_total; nl[a] = nv[a] ? n : 0; }
#pragma unroll
    for (int b = 0; b < MA; ++b) { int m = cm + b * ncm; mv[b] = m < cm_block; ml[b] = mv[b] ? m : 0; }
    float acc[NA][MA][G];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < MA; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) acc[a][b][g] = 0.f;
    const int rbeg = split * rows_per;
#pragma unroll 2
    for (int r = rbeg; r < rbeg + rows_per; ++r) {
      float dv[NA][B], xv[MA][B];
#pragma unroll
      for (int a = 0; a < NA; ++a) load_vec<B>(dv[a], nbuf + r * nstride + nl[a] * B);
#pragma unroll
      for (int b = 0; b < MA; ++b) load_vec<B>(xv[b], mbuf + r * mstride + ml[b] * B);
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < MA; ++b)
#pragma unroll
          for (int i = 0; i < B; ++i) acc[a][b][A::grade_of(i)] = fmaf(dv[a][i], xv[b][i], acc[a][b][A::grade_of(i)]);
    }
    float* out = gacc + (size_t)split * cn_total * m_ld * G;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      if (!nv[a]) continue;
#pragma unroll
      for (int b = 0; b < MA; ++b) {
        if (!mv[b]) continue;
        float* o = out + ((size_t)nl[a] * m_ld + m_off + ml[b]) * G;
        if constexpr (G == 4) {
          float4 t = *reinterpret_cast<float4*>(o);
          t.x += acc[a][b][0]; t.y += acc[a][b][1]; t.z += acc[a][b][2]; t.w += acc[a][b][3];
          *reinterpret_cast<float4*>(o) = t;
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g) o[g] += acc[a][b][g];
        }
      }
    }
  }
}

struct BwdWorkspace {
  float* dw1;    // [grid][ks][c][cin][G]
  float* dwr;    // [grid][ks][c][c][G]
  float* dwl;    // [grid][ks][c][c][G]
  float* small;  // [grid][nwarps][c * (P + 3G + 3)]:  dw[c][P] | dna[c][G] | dsa[c][G] | dsb[c][G] | dla[c] | db1[c] | dbl[c]
};

// one pass of the transposed W1 GEMM over output channels [ob, ob + width) with NCHT channels per thread
template <int DIM, int NCHT>
__device__ __forceinline__ void dx_pass(const float* __restrict__ bufB, const FusedPlan& p, const WRef& w1, float* grad_x,
                                        int64_t row0, int64_t rows, int cin, int ob, int width, int c, int rg, bool active) {
  using Cfg = GemmCfg<DIM>;
  constexpr int B = Alg<DIM>::B, RB = Cfg::RB;
  float acc[RB][NCHT][B];
#pragma unroll
  for (int j = 0; j < RB; ++j)
#pragma unroll
    for (int a = 0; a < NCHT; ++a)
#pragma unroll
      for (int i = 0; i < B; ++i) acc[j][a][i] = 0.f;
  const bool act = active && c < width;
  gemm_run<DIM, true, NCHT>(acc, bufB + rg * RB * p.sb, p.sb, w1, act, c, p.nc, ob, width);
  if (act && grad_x) {
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int64_t r = row0 + rg * RB + j;
      if (r >= rows) continue;
#pragma unroll
      for (int a = 0; a < NCHT; ++a) {
        const int ol = c + a * p.nc;
        if (ol < width) store_vec<B>(grad_x + (r * cin + ob + ol) * B, acc[j][a]);
      }
    }
  }
}

// ===================================================================================================
// backward
template <int DIM>
__global__ void __launch_bounds__(256, 1) block_bwd_kernel(csmpn_block_desc d, csmpn_block_grads gr, FusedPlan p,
                                                           BwdWorkspace ws) {
  using A = Alg<DIM>;
  using Cfg = GemmCfg<DIM>;
  constexpr int B = A::B, G = A::G, P = A::P, RB = Cfg::RB, NCH = Cfg::NCH;
  extern __shared__ __align__(128) float smem[];
  float* bufA = smem;
  float* bufB = bufA + p.tr * p.sa;
  float* wsm = bufB + p.tr * p.sb;
  float* part = wsm + p.wbuf;            // [tr][nc] sum of norms
  float* part2 = part + p.tr * p.nc;     // [tr][nc] sum of a_n <dy, o>
  uint64_t* bar = reinterpret_cast<uint64_t*>(part2 + p.tr * p.nc);
  const int C = d.c, cin = d.c0 + d.c1 + d.c2;
  const int c = threadIdx.x % p.nc, rg = threadIdx.x / p.nc;
  const bool active = rg < p.rg;
  const int warp = threadIdx.x >> 5;
  int och[NCH];
  bool ov[NCH];
#pragma unroll
  for (int a = 0; a < NCH; ++a) { och[a] = c + a * p.nc; ov[a] = active && och[a] < C; if (!ov[a]) och[a] = 0; }
  const int64_t tiles = (d.rows + p.tr - 1) / p.tr;
  const int small_words = C * (P + 3 * G + 3);
  float* my_small = ws.small + ((size_t)blockIdx.x * p.nwarps + warp) * small_words;
  float* s_dw = my_small;
  float* s_dna = s_dw + C * P;
  float* s_dsa = s_dna + C * G;
  float* s_dsb = s_dsa + C * G;
  float* s_dla = s_dsb + C * G;
  float* s_db1 = s_dla + C;
  float* s_dbl = s_db1 + C;
  float* my_dw1 = ws.dw1 + (size_t)blockIdx.x * p.ks * C * cin * G;
  float* my_dwr = ws.dwr + (size_t)blockIdx.x * p.ks * C * C * G;
  float* my_dwl = ws.dwl + (size_t)blockIdx.x * p.ks * C * C * G;
  WRef w1, wr, wl;
  setup_weights<DIM>(d, p, wsm, w1, wr, wl);
  if (!p.resident) { w1.kchunk = p.kt1; wr.kchunk = p.ktc; wl.kchunk = p.ktc; }
  if (threadIdx.x == 0) mbar_init(bar, 32);
  uint32_t phase = 0;

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * p.tr;
    const int lr0 = rg * RB;
    float dd[RB][NCH][B];   // running gradient held by this thread (dy -> d = do / sqrt2)
    float nu_sum[RB], dot_sum[RB];
    // ---- B0: load o, dy; layer-norm statistics
    {
      float oo[RB][NCH][B];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        nu_sum[j] = 0.f; dot_sum[j] = 0.f;
        const int64_t r = row0 + lr0 + j;
#pragma unroll
        for (int a = 0; a < NCH; ++a) {
          if (ov[a] && r < d.rows) {
            load_vec<B>(oo[j][a], d.save_o + (r * C + och[a]) * B);
            load_vec<B>(dd[j][a], gr.grad_y + (r * C + och[a]) * B);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < B; ++i) dot = fmaf(dd[j][a][i], oo[j][a][i], dot);
            nu_sum[j] += smooth_abs_sqrt(mv_sumsq<DIM>(oo[j][a]));
            dot_sum[j] = fmaf(d.la[och[a]], dot, dot_sum[j]);
          } else {
#pragma unroll
            for (int i = 0; i < B; ++i) { oo[j][a][i] = 0.f; dd[j][a][i] = 0.f; }
          }
        }
      }
      __syncthreads();  // previous tile's readers of part / part2 / bufA / bufB are done
      if (active) {
#pragma unroll
        for (int j = 0; j < RB; ++j) { part[(lr0 + j) * p.nc + c] = nu_sum[j]; part2[(lr0 + j) * p.nc + c] = dot_sum[j]; }
      }
      __syncthreads();
      float g_la[NCH][1];
#pragma unroll
      for (int a = 0; a < NCH; ++a) g_la[a][0] = 0.f;
      if (active) {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          float s1 = 0.f, s2 = 0.f;
          for (int cc = 0; cc < p.nc; ++cc) { s1 += part[(lr0 + j) * p.nc + cc]; s2 += part2[(lr0 + j) * p.nc + cc]; }
          const float mu = s1 / (float)C + kEps;
          const float inv_mu = 1.f / mu;
          const float dmu_c = -s2 * inv_mu * inv_mu / (float)C;  // d loss / d mu, divided by C
#pragma unroll
          for (int a = 0; a < NCH; ++a) {
            if (!ov[a]) continue;
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < B; ++i) dot = fmaf(dd[j][a][i], oo[j][a][i], dot);
            g_la[a][0] = fmaf(dot, inv_mu, g_la[a][0]);
            const float Q = mv_sumsq<DIM>(oo[j][a]);
            const float nu = smooth_abs_sqrt(Q);
            const float k1 = d.la[och[a]] * inv_mu * kInvSqrt2;
            const float k2 = dmu_c * Q / (nu * nu * nu) * kInvSqrt2;
            // d = do / sqrt2  (gradient of both the left branch xl and the product z)
#pragma unroll
            for (int i = 0; i < B; ++i) dd[j][a][i] = fmaf(k1, dd[j][a][i], k2 * oo[j][a][i]);
          }
        }
      }
      warp_reduce_store<NCH, 1>(g_la, s_dla, och, ov, p.nc);
    }
    // ---- B1: y2 = silu(y1) -> bufB; product / normalisation backward; d -> bufA
    float dy2[RB][NCH][B];  // gradient w.r.t. y2 (starts with the product's left-operand term)
    float dxr[RB][NCH][B];
    {
      float g_w[NCH][P], g_na[NCH][G], g_bl[NCH][1];
#pragma unroll
      for (int a = 0; a < NCH; ++a) {
        g_bl[a][0] = 0.f;
#pragma unroll
        for (int q = 0; q < P; ++q) g_w[a][q] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) g_na[a][g] = 0.f;
      }
#pragma unroll
      for (int a = 0; a < NCH; ++a) {
        const int n = och[a];
        float sa_[G], sb_[G], s[G], wv[P];
#pragma unroll
        for (int g = 0; g < G; ++g) { sa_[g] = d.sa[n * G + g]; sb_[g] = d.sb[n * G + g]; s[g] = sigmoidf_(d.na[n * G + g]); }
#pragma unroll
        for (int q = 0; q < P; ++q) wv[q] = d.wp[n * P + q];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int64_t r = row0 + lr0 + j;
          float y2[B], xr[B];
          if (ov[a] && r < d.rows) {
            float y1[B], sg[G], inv[G];
            load_vec<B>(y1, d.save_y1 + (r * C + n) * B);
            load_vec<B>(xr, d.save_xr + (r * C + n) * B);
            silu_gates<DIM>(y1, sa_, sb_, sg, inv);
#pragma unroll
            for (int i = 0; i < B; ++i) y2[i] = y1[i] * sg[A::grade_of(i)];
          } else {
#pragma unroll
            for (int i = 0; i < B; ++i) { y2[i] = 0.f; xr[i] = 0.f; }
          }
          if (ov[a]) store_vec<B>(bufB + (lr0 + j) * p.sb + n * B, y2);
          float q[G], nrm[G], rinv[G], xn[B], dxn[B];
          norm_factors<DIM>(xr, s, q, nrm, rinv);
#pragma unroll
          for (int i = 0; i < B; ++i) { xn[i] = xr[i] * rinv[A::grade_of(i)]; dy2[j][a][i] = 0.f; dxn[i] = 0.f; }
          A::template wgp_bwd<false>(y2, xn, wv, dd[j][a], nullptr, dy2[j][a], dxn, g_w[a]);
          // normalisation backward: xn_i = xr_i * rinv_g
          float t[G];
#pragma unroll
          for (int g = 0; g < G; ++g) t[g] = 0.f;
#pragma unroll
          for (int i = 0; i < B; ++i) t[A::grade_of(i)] = fmaf(dxn[i], xr[i], t[A::grade_of(i)]);
          float coef[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float ddn = -t[g] * rinv[g] * rinv[g];  // d loss / d denominator
            g_na[a][g] = fmaf(ddn * (nrm[g] - 1.f), s[g] * (1.f - s[g]), g_na[a][g]);
            coef[g] = ddn * s[g] * q[g] / (nrm[g] * nrm[g] * nrm[g]);
          }
#pragma unroll
          for (int i = 0; i < B; ++i) dxr[j][a][i] = fmaf(dxn[i], rinv[A::grade_of(i)], coef[A::grade_of(i)] * xr[i]);
          g_bl[a][0] += dd[j][a][0];
          if (ov[a]) store_vec<B>(bufA + (lr0 + j) * p.sa + n * B, dd[j][a]);
        }
      }
      warp_reduce_store<NCH, P>(g_w, s_dw, och, ov, p.nc);
      warp_reduce_store<NCH, G>(g_na, s_dna, och, ov, p.nc);
      warp_reduce_store<NCH, 1>(g_bl, s_dbl, och, ov, p.nc);
    }
    // ---- B2: dy2 += D * WL ; dWL += D^T Y2
    gemm_run<DIM, true, NCH>(dy2, bufA + lr0 * p.sa, p.sa, wl, active, c, p.nc, 0, C);
    dw_tile<DIM>(bufA, p.sa, C, bufB, p.sb, C, 0, C, p.tr, p.ks, my_dwl);
    __syncthreads();  // all readers of D (bufA) are done
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
#pragma unroll
      for (int j = 0; j < RB; ++j) store_vec<B>(bufA + (lr0 + j) * p.sa + och[a] * B, dxr[j][a]);
    }
    // ---- B3: dy2 += DXR * WR ; dWR += DXR^T Y2
    gemm_run<DIM, true, NCH>(dy2, bufA + lr0 * p.sa, p.sa, wr, active, c, p.nc, 0, C);
    dw_tile<DIM>(bufA, p.sa, C, bufB, p.sb, C, 0, C, p.tr, p.ks, my_dwr);
    // ---- B4: MVSiLU backward -> dy1 (registers, reuse dy2)
    {
      float g_sa[NCH][G], g_sb[NCH][G], g_b1[NCH][1];
#pragma unroll
      for (int a = 0; a < NCH; ++a) {
        g_b1[a][0] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) { g_sa[a][g] = 0.f; g_sb[a][g] = 0.f; }
      }
#pragma unroll
      for (int a = 0; a < NCH; ++a) {
        const int n = och[a];
        float sa_[G], sb_[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { sa_[g] = d.sa[n * G + g]; sb_[g] = d.sb[n * G + g]; }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int64_t r = row0 + lr0 + j;
          if (!(ov[a] && r < d.rows)) {
#pragma unroll
            for (int i = 0; i < B; ++i) dy2[j][a][i] = 0.f;
            continue;
          }
          float y1[B], sg[G], inv[G], t[G];
          load_vec<B>(y1, d.save_y1 + (r * C + n) * B);
          silu_gates<DIM>(y1, sa_, sb_, sg, inv);
#pragma unroll
          for (int g = 0; g < G; ++g) t[g] = 0.f;
#pragma unroll
          for (int i = 0; i < B; ++i) t[A::grade_of(i)] = fmaf(dy2[j][a][i], y1[i], t[A::grade_of(i)]);
          float ds[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            ds[g] = t[g] * sg[g] * (1.f - sg[g]);
            g_sa[a][g] = fmaf(ds[g], inv[g], g_sa[a][g]);
            g_sb[a][g] += ds[g];
          }
#pragma unroll
          for (int i = 0; i < B; ++i) {
            const int g = A::grade_of(i);
            const float dinv = (g == 0) ? 1.f : 2.f * y1[i];
            dy2[j][a][i] = fmaf(sg[g], dy2[j][a][i], ds[g] * sa_[g] * dinv);
          }
          g_b1[a][0] += dy2[j][a][0];
        }
      }
      warp_reduce_store<NCH, G>(g_sa, s_dsa, och, ov, p.nc);
      warp_reduce_store<NCH, G>(g_sb, s_dsb, och, ov, p.nc);
      warp_reduce_store<NCH, 1>(g_b1, s_db1, och, ov, p.nc);
    }
    // ---- B4.5: the input rows again (TMA, L2-resident) for dW1; then dy1 -> bufB
    fence_proxy_async();
    __syncthreads();  // generic readers of bufB (Y2) and bufA (DXR) are done
    load_input_tile<DIM>(bufA, p.sa, bufB, p.sb, d, row0, p.tr, bar, phase);
    phase ^= 1;
    __syncthreads();  // the input pass has finished reading the sender rows in bufB
#pragma unroll
    for (int a = 0; a < NCH; ++a) {
      if (!ov[a]) continue;
#pragma unroll
      for (int j = 0; j < RB; ++j) store_vec<B>(bufB + (lr0 + j) * p.sb + och[a] * B, dy2[j][a]);
    }
    // ---- B5: dx0 = DY1 * W1 in column blocks of C channels (written straight to global);  dW1 += DY1^T X0
    for (int ob = 0; ob < cin; ob += C) {
      const int width = (cin - ob) < C ? (cin - ob) : C;
      const int ncht = (width + p.nc - 1) / p.nc;  // channels per thread in this block
      if (ncht >= NCH) dx_pass<DIM, NCH>(bufB, p, w1, gr.grad_x, row0, d.rows, cin, ob, width, c, rg, active);
      else if (ncht == 1) dx_pass<DIM, 1>(bufB, p, w1, gr.grad_x, row0, d.rows, cin, ob, width, c, rg, active);
      else if (ncht == 2) dx_pass<DIM, (NCH > 2 ? 2 : NCH)>(bufB, p, w1, gr.grad_x, row0, d.rows, cin, ob, width, c, rg, active);
      else dx_pass<DIM, (NCH > 3 ? 3 : NCH)>(bufB, p, w1, gr.grad_x, row0, d.rows, cin, ob, width, c, rg, active);
      dw_tile<DIM>(bufB, p.sb, C, bufA + ob * B, p.sa, width, ob, cin, p.tr, p.ks, my_dw1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// final fixed-order reduction of the per-CTA partials into the parameter gradients: one warp per element,
// lane l sums parts l, l+32, ...; fixed xor tree across lanes.
struct FinalSeg { const float* in; float* out; int n; int parts; int64_t stride; };
struct FinalSegs { FinalSeg s[10]; int count; int total; };

__global__ void __launch_bounds__(256) block_bwd_final_kernel(FinalSegs segs) {
  int q = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= segs.total) return;
  int k = 0;
  while (k < segs.count - 1 && q >= segs.s[k].n) { q -= segs.s[k].n; ++k; }
  const FinalSeg& sg = segs.s[k];
  float s = 0.f;
  for (int pidx = lane; pidx < sg.parts; pidx += 32) s += sg.in[(size_t)pidx * sg.stride + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0 && sg.out != nullptr && q < sg.n) sg.out[q] = s;
}

template <int DIM>
int launch_block_fwd(const csmpn_block_desc& d, cudaStream_t s) {
  FusedPlan p;
  int st = make_fused_plan<DIM>(d, &p);
  if (st) return st;
  static bool attr_set = false;
  if (!attr_set) {
    CSMPN_CUDA_TRY(cudaFuncSetAttribute(block_fwd_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem + 1024));
    attr_set = true;
  }
  block_fwd_kernel<DIM><<<p.grid, p.threads, p.smem, s>>>(d, p);
  CSMPN_LAUNCH_CHECK("block_fwd");
  return CSMPN_OK;
}

template <int DIM>
int64_t block_bwd_ws_bytes(const csmpn_block_desc& d, FusedPlan* pp) {
  FusedPlan p;
  if (make_fused_plan<DIM>(d, &p)) return -1;
  if (pp) *pp = p;
  constexpr int G = Alg<DIM>::G, P = Alg<DIM>::P;
  const int64_t c = d.c, cin = d.c0 + d.c1 + d.c2;
  int64_t words = (int64_t)p.grid * ((int64_t)p.ks * (c * cin * G + 2 * c * c * G) + (int64_t)p.nwarps * c * (P + 3 * G + 3));
  return words * 4;
}

template <int DIM>
int launch_block_bwd(const csmpn_block_desc& d, const csmpn_block_grads& g, void* workspace, int64_t bytes, cudaStream_t s) {
  FusedPlan p;
  const int64_t need = block_bwd_ws_bytes<DIM>(d, &p);
  if (need < 0) return CSMPN_ERR_UNSUPPORTED;
  if (!workspace || bytes < need) return CSMPN_ERR_WORKSPACE;
  constexpr int G = Alg<DIM>::G, P = Alg<DIM>::P;
  const int c = d.c, cin = d.c0 + d.c1 + d.c2;
  BwdWorkspace ws;
  ws.dw1 = (float*)workspace;
  ws.dwr = ws.dw1 + (size_t)p.grid * p.ks * c * cin * G;
  ws.dwl = ws.dwr + (size_t)p.grid * p.ks * c * c * G;
  ws.small = ws.dwl + (size_t)p.grid * p.ks * c * c * G;
  CSMPN_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)need, s));
  static bool attr_set = false;
  if (!attr_set) {
    CSMPN_CUDA_TRY(cudaFuncSetAttribute(block_bwd_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem + 1024));
    attr_set = true;
  }
  block_bwd_kernel<DIM><<<p.grid, p.threads, p.smem, s>>>(d, g, p, ws);
  CSMPN_LAUNCH_CHECK("block_bwd");
  const int small_words = c * (P + 3 * G + 3);
  const int sparts = p.grid * p.nwarps;
  FinalSegs fs;
  int k = 0;
  auto add = [&](const float* in, float* out, int n, int parts, int64_t stride) {
    fs.s[k].in = in; fs.s[k].out = out; fs.s[k].n = n; fs.s[k].parts = parts; fs.s[k].stride = stride; ++k;
  };
  add(ws.dw1, g.g_w1, c * cin * G, p.grid * p.ks, (int64_t)c * cin * G);
  add(ws.dwr, g.g_wr, c * c * G, p.grid * p.ks, (int64_t)c * c * G);
  add(ws.dwl, g.g_wl, c * c * G, p.grid * p.ks, (int64_t)c * c * G);
  const float* sm = ws.small;
  add(sm, g.g_wp, c * P, sparts, small_words); sm += c * P;
  add(sm, g.g_na, c * G, sparts, small_words); sm += c * G;
  add(sm, g.g_sa, c * G, sparts, small_words); sm += c * G;
  add(sm, g.g_sb, c * G, sparts, small_words); sm += c * G;
  add(sm, g.g_la, c, sparts, small_words); sm += c;
  add(sm, d.has_b1 ? g.g_b1 : nullptr, c, sparts, small_words); sm += c;
  add(sm, g.g_bl, c, sparts, small_words);
  fs.count = k;
  fs.total = 0;
  for (int i = 0; i < k; ++i) fs.total += fs.s[i].n;
  const int64_t threads = (int64_t)fs.total * 32;
  block_bwd_final_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(fs);
  CSMPN_LAUNCH_CHECK("block_bwd_final");
  return CSMPN_OK;
}

inline int check_desc(int dim, const csmpn_block_desc* d) {
  if (!d) return CSMPN_ERR_BAD_ARG;
  if (dim != 2 && dim != 3 && dim != 5) return CSMPN_ERR_UNSUPPORTED;
  if (d->rows < 0 || d->c <= 0 || d->c0 <= 0 || d->c1 < 0 || d->c2 < 0) return CSMPN_ERR_BAD_ARG;
  if (d->mode != 0 && d->mode != 1) return CSMPN_ERR_BAD_ARG;
  if (!d->p0 || (d->c1 > 0 && !d->p1) || (d->c2 > 0 && !d->p2)) return CSMPN_ERR_BAD_ARG;
  if (d->mode == 1 && (!d->src || !d->dst || (d->c1 > 0 && !d->eid) || d->c2 != 0)) return CSMPN_ERR_BAD_ARG;
  if (!d->w1 || !d->sa || !d->sb || !d->wr || !d->na || !d->wl || !d->bl || !d->wp || !d->la) return CSMPN_ERR_BAD_ARG;
  if (d->has_b1 && !d->b1) return CSMPN_ERR_BAD_ARG;
  return CSMPN_OK;
}

}  // namespace csmpn

using namespace csmpn;

extern "C" {

int csmpn_block_fwd(int dim, const csmpn_block_desc* desc, csmpn_stream_t stream) {
  int st = check_desc(dim, desc);
  if (st) return st;
  if (!desc->y) return CSMPN_ERR_BAD_ARG;
  if (desc->rows == 0) return CSMPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim) {
    case 2: return launch_block_fwd<2>(*desc, s);
    case 3: return launch_block_fwd<3>(*desc, s);
    case 5: return launch_block_fwd<5>(*desc, s);
  }
  return CSMPN_ERR_UNSUPPORTED;
}

int64_t csmpn_block_bwd_workspace(int dim, const csmpn_block_desc* desc) {
  if (check_desc(dim, desc)) return -1;
  switch (dim) {
    case 2: return block_bwd_ws_bytes<2>(*desc, nullptr);
    case 3: return block_bwd_ws_bytes<3>(*desc, nullptr);
    case 5: return block_bwd_ws_bytes<5>(*desc, nullptr);
  }
  return -1;
}

int csmpn_block_bwd(int dim, const csmpn_block_desc* desc, const csmpn_block_grads* grads, void* workspace,
                    int64_t workspace_bytes, csmpn_stream_t stream) {
  int st = check_desc(dim, desc);
  if (st) return st;
  if (!grads || !grads->grad_y || !desc->save_y1 || !desc->save_xr || !desc->save_o) return CSMPN_ERR_BAD_ARG;
  if (!grads->g_w1 || !grads->g_sa || !grads->g_sb || !grads->g_wr || !grads->g_na || !grads->g_wl || !grads->g_bl ||
      !grads->g_wp || !grads->g_la)
    return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim) {
    case 2: return launch_block_bwd<2>(*desc, *grads, workspace, workspace_bytes, s);
    case 3: return launch_block_bwd<3>(*desc, *grads, workspace, workspace_bytes, s);
    case 5: return launch_block_bwd<5>(*desc, *grads, workspace, workspace_bytes, s);
  }
  return CSMPN_ERR_UNSUPPORTED;
}

}  // extern "C"
