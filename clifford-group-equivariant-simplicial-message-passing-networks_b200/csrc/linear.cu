// MVLinear forward / input-gradient / weight-gradient kernels (reference: csmpn/models/cegnn_utils.py:287-338).
#include "gemm.cuh"

namespace csmpn {

struct LinearPlan {
  int tr;       // rows per tile
  int kc;       // reduction-channel chunk held in shared memory
  int nc;       // output channel groups (threads along channels)
  int threads;
  int sx, sw;   // shared-memory row strides (words)
  size_t smem;
  int grid;
};

template <int DIM>
inline LinearPlan make_linear_plan(int64_t rows, int kdim, int odim, bool trans) {
  using Cfg = GemmCfg<DIM>;
  constexpr int B = Alg<DIM>::B;
  LinearPlan p;
  p.nc = (odim + Cfg::NCH - 1) / Cfg::NCH;
  int rgroups = 256 / p.nc;
  if (rgroups < 1) rgroups = 1;
  int max_rg = (DIM <= 3 ? 32 : 16) / Cfg::RB;
  if (rgroups > max_rg) rgroups = max_rg;
  p.tr = rgroups * Cfg::RB;
  p.threads = ((rgroups * p.nc + 31) / 32) * 32;
  int kc_w = 16384 / (odim * Cfg::GP);
  int kc_x = 12288 / (p.tr * B);
  int kc = kc_w < kc_x ? kc_w : kc_x;
  if (kc < 4) kc = 4;
  kc &= ~3;
  if (kc > kdim) kc = kdim;
  p.kc = kc;
  p.sx = pad_stride(kc * B);
  p.sw = trans ? pad_stride(odim * Cfg::GP) : pad_stride(kc * Cfg::GP);
  size_t wsz = trans ? (size_t)kc * p.sw : (size_t)odim * p.sw;
  p.smem = ((size_t)p.tr * p.sx + wsz) * sizeof(float);
  int64_t tiles = (rows + p.tr - 1) / p.tr;
  int64_t cap = (int64_t)sm_count_cached() * 2;
  p.grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
  return p;
}

// y[r,o,:] = sum_k x[r,k,:] * W(k,o)[grade]   (+ bias[o] on blade 0)
template <int DIM, bool TRANS>
__global__ void __launch_bounds__(256) mvlinear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y,
                                                       int64_t rows, int c_in, int c_out, int gw, LinearPlan p) {
  using A = Alg<DIM>;
  using Cfg = GemmCfg<DIM>;
  constexpr int B = A::B, RB = Cfg::RB, NCH = Cfg::NCH, GP = Cfg::GP;
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* ws = smem + p.tr * p.sx;
  const int kdim = TRANS ? c_out : c_in;
  const int odim = TRANS ? c_in : c_out;
  const int c = threadIdx.x % p.nc, rg = threadIdx.x / p.nc;
  const bool active = rg * RB < p.tr;
  const int nchunks = (kdim + p.kc - 1) / p.kc;

  const float* wp[NCH];
  int och[NCH];
#pragma unroll
  for (int a = 0; a < NCH; ++a) {
    int o = c + a * p.nc;
    och[a] = o;
    int oc = o < odim ? o : 0;
    wp[a] = TRANS ? ws + oc * GP : ws + oc * p.sw;
  }
  const int wk_stride = TRANS ? p.sw : GP;
  const int64_t tiles = (rows + p.tr - 1) / p.tr;

  if (nchunks == 1) stage_weights<DIM, TRANS>(ws, p.sw, w, c_out, c_in, gw, 0, kdim);

  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * p.tr;
    float acc[RB][NCH][B];
#pragma unroll
    for (int j = 0; j < RB; ++j)
#pragma unroll
      for (int a = 0; a < NCH; ++a)
#pragma unroll
        for (int i = 0; i < B; ++i) acc[j][a][i] = 0.f;

    for (int ch = 0; ch < nchunks; ++ch) {
      const int k0 = ch * p.kc;
      const int kc = (kdim - k0) < p.kc ? (kdim - k0) : p.kc;
      __syncthreads();  // previous consumers of xs / ws are done
      stage_rows<DIM>(xs, p.sx, x, row0, rows, p.tr, kdim, k0, kc);
      if (nchunks > 1) stage_weights<DIM, TRANS>(ws, p.sw, w, c_out, c_in, gw, k0, kc);
      __syncthreads();
      if (active) gemm_accumulate<DIM>(acc, xs + rg * RB * p.sx, p.sx, wp, wk_stride, kc);
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int64_t r = row0 + rg * RB + j;
        if (r >= rows) continue;
#pragma unroll
        for (int a = 0; a < NCH; ++a) {
          if (och[a] >= odim) continue;
          if (bias != nullptr) acc[j][a][0] += bias[och[a]];
          store_vec<B>(y + (r * odim + och[a]) * B, acc[j][a]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: dW[n,m,g] = sum_r sum_{i in g} dy[r,n,i] x[r,m,i];  db[n] = sum_r dy[r,n,0]
template <int DIM> struct DwCfg { static constexpr int NA = (Alg<DIM>::G <= 4) ? 4 : 2, MA = 4; };
constexpr int kDwTileN = 64, kDwTileM = 64, kDwRows = 16;

struct DwPlan {
  int tiles_n, tiles_m, splits, ncn, ncm, threads, sy, sxw;
  int64_t rows_per_split;
  size_t smem;
};

template <int DIM>
inline DwPlan make_dw_plan(int64_t rows, int c_in, int c_out) {
  using D = DwCfg<DIM>;
  constexpr int B = Alg<DIM>::B;
  DwPlan p;
  p.tiles_n = (c_out + kDwTileN - 1) / kDwTileN;
  p.tiles_m = (c_in + kDwTileM - 1) / kDwTileM;
  int tn = c_out < kDwTileN ? c_out : kDwTileN, tm = c_in < kDwTileM ? c_in : kDwTileM;
  p.ncn = (tn + D::NA - 1) / D::NA;
  p.ncm = (tm + D::MA - 1) / D::MA;
  p.threads = ((p.ncn * p.ncm + 31) / 32) * 32;
  p.sy = pad_stride(tn * B);
  p.sxw = pad_stride(tm * B);
  p.smem = (size_t)kDwRows * (p.sy + p.sxw) * sizeof(float);
  int64_t row_tiles = (rows + kDwRows - 1) / kDwRows;
  int64_t cap = (int64_t)sm_count_cached() * 4 / (p.tiles_n * p.tiles_m);
  if (cap < 1) cap = 1;
  if (cap > 512) cap = 512;
  int64_t splits = row_tiles < cap ? row_tiles : cap;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  int64_t rt_per = (row_tiles + splits - 1) / splits;
  p.rows_per_split = rt_per * kDwRows;
  return p;
}

template <int DIM>
__global__ void __launch_bounds__(256) mvlinear_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ ws_out, float* __restrict__ ws_bias,
                                                          int64_t rows, int c_in, int c_out, DwPlan p) {
  using A = Alg<DIM>;
  using D = DwCfg<DIM>;
  constexpr int B = A::B, G = A::G, NA = D::NA, MA = D::MA;
  extern __shared__ __align__(16) float smem[];
  float* dys = smem;
  float* xs = smem + kDwRows * p.sy;
  const int split = blockIdx.x;
  const int tn = blockIdx.y / p.tiles_m, tmi = blockIdx.y % p.tiles_m;
  const int n_base = tn * kDwTileN, m_base = tmi * kDwTileM;
  const int tile_n = (c_out - n_base) < kDwTileN ? (c_out - n_base) : kDwTileN;
  const int tile_m = (c_in - m_base) < kDwTileM ? (c_in - m_base) : kDwTileM;
  const int cm = threadIdx.x % p.ncm, cn = threadIdx.x / p.ncm;
  const bool active = cn < p.ncn;
  int nl[NA], ml[MA];  // local channel indices inside the tile (clamped for loads)
  bool nv[NA], mv[MA];
#pragma unroll
  for (int a = 0; a < NA; ++a) { int n = cn + a * p.ncn; nv[a] = active && n < tile_n; nl[a] = nv[a] ? n : 0; }
#pragma unroll
  for (int b = 0; b < MA; ++b) { int m = cm + b * p.ncm; mv[b] = active && m < tile_m; ml[b] = mv[b] ? m : 0; }

  float acc[NA][MA][G];
  float accb[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    accb[a] = 0.f;
#pragma unroll
    for (int b = 0; b < MA; ++b)
#pragma unroll
      for (int g = 0; g < G; ++g) acc[a][b][g] = 0.f;
  }

  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  int64_t r_end = r_begin + p.rows_per_split;
  if (r_end > rows) r_end = rows;
  constexpr int V = (B % 4 == 0) ? 4 : 2;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kDwRows) {
    __syncthreads();
    {  // stage dy[:, n_base : n_base+tile_n, :] and x[:, m_base : m_base+tile_m, :]
      const int vy = tile_n * B / V, vx = tile_m * B / V;
      for (int idx = threadIdx.x; idx < kDwRows * (vy + vx); idx += blockDim.x) {
        const int r = idx / (vy + vx), v = idx - r * (vy + vx);
        const int64_t gr = r0 + r;
        const bool ok = gr < r_end;
        const float* src = v < vy ? dy + (gr * c_out + n_base) * B + V * v : x + (gr * c_in + m_base) * B + V * (v - vy);
        float* dst = v < vy ? dys + r * p.sy + V * v : xs + r * p.sxw + V * (v - vy);
        if constexpr (V == 4) *reinterpret_cast<float4*>(dst) = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0, 0, 0, 0);
        else *reinterpret_cast<float2*>(dst) = ok ? *reinterpret_cast<const float2*>(src) : make_float2(0, 0);
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll 2
      for (int r = 0; r < kDwRows; ++r) {
        float dv[NA][B], xv[MA][B];
#pragma unroll
        for (int a = 0; a < NA; ++a) load_vec<B>(dv[a], dys + r * p.sy + nl[a] * B);
#pragma unroll
        for (int b = 0; b < MA; ++b) load_vec<B>(xv[b], xs + r * p.sxw + ml[b] * B);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          accb[a] += dv[a][0];
#pragma unroll
          for (int b = 0; b < MA; ++b)
#pragma unroll
            for (int i = 0; i < B; ++i) acc[a][b][A::grade_of(i)] = fmaf(dv[a][i], xv[b][i], acc[a][b][A::grade_of(i)]);
        }
      }
    }
  }
  // partial: ws_out[split][n][m][g]
  float* out = ws_out + (size_t)split * c_out * c_in * G;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    if (!nv[a]) continue;
    const int n = n_base + nl[a];
#pragma unroll
    for (int b = 0; b < MA; ++b) {
      if (!mv[b]) continue;
      const int m = m_base + ml[b];
#pragma unroll
      for (int g = 0; g < G; ++g) out[((size_t)n * c_in + m) * G + g] = acc[a][b][g];
    }
    if (tmi == 0 && cm == 0 && ws_bias) ws_bias[(size_t)split * c_out + n] = accb[a];
  }
}

// grad_w = sum over splits (fixed order); gw == 1 additionally sums the grades (subspaces=False).
// 16 consecutive outputs per CTA x 16 split groups: thread (e, pg) sums splits pg, pg+16, ... (independent loads), then
// the 16 group sums are combined in a fixed order.
__global__ void __launch_bounds__(256) dw_final_kernel(const float* __restrict__ ws_out, const float* __restrict__ ws_bias,
                                                       float* __restrict__ grad_w, float* __restrict__ grad_bias, int c_out,
                                                       int c_in, int G, int gw, int splits) {
  __shared__ float red[16][17];
  const int e = threadIdx.x & 15, pg = threadIdx.x >> 4;
  const int total_w = c_out * c_in * gw;
  const int q = blockIdx.x * 16 + e;
  float s = 0.f;
  float* outp = nullptr;
  if (q < total_w) {
    if (gw == 1) {
      for (int sp = pg; sp < splits; sp += 16) {
        const float* p = ws_out + ((size_t)sp * c_out * c_in + q) * G;
        for (int g = 0; g < G; ++g) s += p[g];
      }
    } else {
#pragma unroll 4
      for (int sp = pg; sp < splits; sp += 16) s += ws_out[(size_t)sp * c_out * c_in * G + q];
    }
    outp = grad_w + q;
  } else if (grad_bias && q < total_w + c_out) {
    const int n = q - total_w;
    for (int sp = pg; sp < splits; sp += 16) s += ws_bias[(size_t)sp * c_out + n];
    outp = grad_bias + n;
  }
  red[pg][e] = s;
  __syncthreads();
  if (pg == 0 && outp) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][e];
    *outp = t;
  }
}

template <int DIM, bool TRANS>
int launch_linear(const float* x, const float* w, const float* bias, float* y, int64_t rows, int c_in, int c_out,
                  int subspaces, cudaStream_t s) {
  const int kdim = TRANS ? c_out : c_in, odim = TRANS ? c_in : c_out;
  LinearPlan p = make_linear_plan<DIM>(rows, kdim, odim, TRANS);
  static bool attr_set = false;
  if (!attr_set) {
    CSMPN_CUDA_TRY(cudaFuncSetAttribute(mvlinear_kernel<DIM, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (p.smem > 200 * 1024) return CSMPN_ERR_UNSUPPORTED;
  mvlinear_kernel<DIM, TRANS><<<p.grid, p.threads, p.smem, s>>>(x, w, bias, y, rows, c_in, c_out,
                                                                  subspaces ? Alg<DIM>::G : 1, p);
  CSMPN_LAUNCH_CHECK("mvlinear");
  return CSMPN_OK;
}

}  // namespace csmpn

using namespace csmpn;

extern "C" {

int csmpn_mvlinear_fwd(int dim, const float* x, const float* weight, const float* bias, float* y, int64_t rows,
                       int c_in, int c_out, int subspaces, csmpn_stream_t stream) {
  if (rows < 0 || c_in <= 0 || c_out <= 0 || !weight) return CSMPN_ERR_BAD_ARG;
  if (rows == 0) return CSMPN_OK;
  if (!x || !y) return CSMPN_ERR_BAD_ARG;
  if (c_out > 1024 || c_in > 4096) return CSMPN_ERR_UNSUPPORTED;
  CSMPN_DISPATCH_DIM(dim, D, {
    return launch_linear<D, false>(x, weight, bias, y, rows, c_in, c_out, subspaces, (cudaStream_t)stream);
  });
  return CSMPN_OK;
}

int csmpn_mvlinear_bwd_input(int dim, const float* grad_y, const float* weight, float* grad_x, int64_t rows, int c_in,
                             int c_out, int subspaces, csmpn_stream_t stream) {
  if (rows < 0 || c_in <= 0 || c_out <= 0 || !weight) return CSMPN_ERR_BAD_ARG;
  if (rows == 0) return CSMPN_OK;
  if (!grad_y || !grad_x) return CSMPN_ERR_BAD_ARG;
  if (c_in > 1024 || c_out > 4096) return CSMPN_ERR_UNSUPPORTED;
  CSMPN_DISPATCH_DIM(dim, D, {
    return launch_linear<D, true>(grad_y, weight, nullptr, grad_x, rows, c_in, c_out, subspaces, (cudaStream_t)stream);
  });
  return CSMPN_OK;
}

int64_t csmpn_mvlinear_bwd_weight_workspace(int dim, int64_t rows, int c_in, int c_out) {
  if (dim < 1 || dim > 5 || c_in <= 0 || c_out <= 0) return 0;
  (void)rows;
  // splits <= 512
  return (int64_t)512 * ((int64_t)c_out * c_in * (dim + 1) + c_out) * (int64_t)sizeof(float);
}

int csmpn_mvlinear_bwd_weight(int dim, const float* x, const float* grad_y, float* grad_w, float* grad_bias,
                              int64_t rows, int c_in, int c_out, int subspaces, void* workspace,
                              int64_t workspace_bytes, csmpn_stream_t stream) {
  if (rows < 0 || c_in <= 0 || c_out <= 0 || !grad_w) return CSMPN_ERR_BAD_ARG;
  if (rows > 0 && (!x || !grad_y)) return CSMPN_ERR_BAD_ARG;
  if (dim < 1 || dim > 5) return CSMPN_ERR_BAD_DIM;
  cudaStream_t s = (cudaStream_t)stream;
  const int G = dim + 1, gw = subspaces ? G : 1;
  CSMPN_DISPATCH_DIM(dim, D, {
    DwPlan p = make_dw_plan<D>(rows, c_in, c_out);
    const int64_t need = (int64_t)p.splits * ((int64_t)c_out * c_in * G + c_out) * 4;
    if (!workspace || workspace_bytes < need) return CSMPN_ERR_WORKSPACE;
    float* ws_out = (float*)workspace;
    float* ws_bias = ws_out + (size_t)p.splits * c_out * c_in * G;
    static bool attr_set = false;
    if (!attr_set) {
      CSMPN_CUDA_TRY(cudaFuncSetAttribute(mvlinear_dw_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    dim3 grid(p.splits, p.tiles_n * p.tiles_m);
    mvlinear_dw_kernel<D><<<grid, p.threads, p.smem, s>>>(x, grad_y, ws_out, ws_bias, rows, c_in, c_out, p);
    CSMPN_LAUNCH_CHECK("mvlinear_dw");
    const int total = c_out * c_in * gw + c_out;
    dw_final_kernel<<<(total + 15) / 16, 256, 0, s>>>(ws_out, ws_bias, grad_w, grad_bias, c_out, c_in, G, gw, p.splits);
    CSMPN_LAUNCH_CHECK("mvlinear_dw_final");
  });
  return CSMPN_OK;
}

}  // extern "C"
