// Graph plumbing of the shared simplicial message layer: CSR construction (stable by construction of the
// final per-segment ordering), gather/difference, deterministic segment reduce and their adjoints.
// Restates what PyG 2.3.0 MessagePassing.propagate + torch_scatter do for EGCL (cegnn_utils.py:277-284):
// gather h_j = h[edge_index[0]], h_i = h[edge_index[1]]; aggregate at edge_index[1] with sum | mean.
// Unlike torch_scatter's atomicAdd the reduction order is fixed: ascending original pair id per receiver.
#include "common.cuh"

namespace csmpn {

// ---------------------------------------------------------------------------------------------------
__global__ void csr_count_kernel(const int64_t* __restrict__ keys, int64_t n_pairs, int64_t n_nodes,
                                 int32_t* __restrict__ counts, int32_t* __restrict__ bad) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pairs; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = keys[e];
    if (k < 0 || k >= n_nodes) { *bad = 1; continue; }
    atomicAdd(&counts[k], 1);
  }
}

constexpr int kScanThreads = 512, kScanPer = 8, kScanTile = kScanThreads * kScanPer;

// exclusive scan of `in[0..n)` into out[0..n], out[n] = total.  Phase 1: per-tile local scan + tile sums.
__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                                  int32_t* __restrict__ tile_sums, int64_t n) {
  __shared__ int32_t warp_sums[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanPer;
  int32_t v[kScanPer], run = 0;
#pragma unroll
  for (int i = 0; i < kScanPer; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    int32_t t = v[i]; v[i] = run; run += t;
  }
  // warp inclusive scan of `run`
  int32_t inc = run;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;  // exclusive
    if (lane == kScanThreads / 32 - 1) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  const int32_t off = warp_sums[wid] + inc - run;
#pragma unroll
  for (int i = 0; i < kScanPer; ++i)
    if (base + i < n) out[base + i] = v[i] + off;
}

// Phase 2: one CTA turns tile sums into exclusive tile offsets (sequential carry over chunks of blockDim).
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ tile_sums, int64_t n_tiles,
                                                         int32_t* __restrict__ total_out) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    int32_t v = i < n_tiles ? tile_sums[i] : 0, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int32_t w = warp_sums[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t excl = carry + warp_sums[wid] + inc - v;
    if (i < n_tiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

// Phase 3: add tile offsets; also seed the fill cursors and write out[n] = total.
__global__ void scan_apply_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ tile_offs, int64_t n,
                                  int32_t* __restrict__ cursor, const int32_t* __restrict__ total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == n) { out[n] = *total; continue; }
    int32_t v = out[i] + tile_offs[i / kScanTile];
    out[i] = v;
    cursor[i] = v;
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ keys, int64_t n_pairs, int64_t n_nodes,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pairs; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = keys[e];
    if (k < 0 || k >= n_nodes) continue;
    int32_t pos = atomicAdd(&cursor[k], 1);
    perm[pos] = (int32_t)e;
  }
}

// Restore the original relative order inside each segment (makes the result independent of atomic order).
__global__ void csr_sort_segments_kernel(const int32_t* __restrict__ rowptr, int64_t n_nodes, int32_t* __restrict__ perm) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = rowptr[n], e = rowptr[n + 1], d = e - b;
    int32_t* a = perm + b;
    if (d <= 1) continue;
    if (d <= 24) {
      for (int i = 1; i < d; ++i) {
        int32_t v = a[i]; int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
      }
    } else {  // heapsort
      for (int start = d / 2 - 1; start >= 0; --start) {
        int root = start; int32_t v = a[root];
        while (true) {
          int child = 2 * root + 1; if (child >= d) break;
          if (child + 1 < d && a[child + 1] > a[child]) ++child;
          if (a[child] <= v) break;
          a[root] = a[child]; root = child;
        }
        a[root] = v;
      }
      for (int end = d - 1; end > 0; --end) {
        int32_t v = a[end]; a[end] = a[0];
        int root = 0;
        while (true) {
          int child = 2 * root + 1; if (child >= end) break;
          if (child + 1 < end && a[child + 1] > a[child]) ++child;
          if (a[child] <= v) break;
          a[root] = a[child]; root = child;
        }
        a[root] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Both CSRs of a batch (by receiver and by sender) + the receiver-sorted views in SIX launches instead of sixteen: every
// kernel handles both key arrays (blockIdx.y), the three-step scan collapses into one CTA per array while the counters fit
// its registers (n_nodes <= 64 k: every configuration BASELINE.json names), and the sorted views and the inverse
// permutation come from one kernel.  The chain is latency-bound (a few microseconds per launch even replayed from a CUDA
// graph) and every host-fed step pays for it.  (A one-CTA-per-array variant with shared-memory counters was slower than
// the sixteen launches: 103 vs 65 us at md17 size, 1.0 ms at NBA size -- profiles/r02_csr_rebuild.log.)
struct CsrPair {
  const int64_t* keys[2];   // [0] receivers (edge_index[1]), [1] senders (edge_index[0])
  int32_t* rowptr[2];
  int32_t* perm[2];
  int32_t* cursor[2];       // workspace: counts, then running cursors
};

__global__ void csr_count2_kernel(CsrPair a, int64_t n_pairs, int64_t n_nodes) {
  const int w = blockIdx.y;
  const int64_t* __restrict__ keys = a.keys[w];
  int32_t* __restrict__ counts = a.cursor[w];
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pairs; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = keys[e];
    if (k >= 0 && k < n_nodes) atomicAdd(&counts[k], 1);
  }
}

constexpr int kScan1Threads = 1024, kScan1Per = 64;  // one CTA scans up to 64 k counters (a contiguous chunk per thread)
__global__ void __launch_bounds__(kScan1Threads) csr_scan2_kernel(CsrPair a, int n_nodes) {
  __shared__ int32_t warp_sums[kScan1Threads / 32];
  const int w = blockIdx.x;
  int32_t* __restrict__ cursor = a.cursor[w];
  int32_t* __restrict__ rowptr = a.rowptr[w];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n_nodes + kScan1Threads - 1) / kScan1Threads;  // <= kScan1Per (host)
  const int b0 = min(tid * per, n_nodes), b1 = min(b0 + per, n_nodes);
  int32_t run = 0;
#pragma unroll 8
  for (int i = b0; i < b1; ++i) run += cursor[i];
  int32_t inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int32_t ws = warp_sums[lane];
    int32_t wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    warp_sums[lane] = wi - ws;
  }
  __syncthreads();
  int32_t off = warp_sums[wid] + inc - run;
#pragma unroll 8
  for (int i = b0; i < b1; ++i) {  // second read of the thread's own chunk: L1 hits
    const int32_t c = cursor[i];
    rowptr[i] = off; cursor[i] = off;
    off += c;
  }
  if (tid == kScan1Threads - 1) rowptr[n_nodes] = off;
}

__global__ void csr_fill2_kernel(CsrPair a, int64_t n_pairs, int64_t n_nodes) {
  const int w = blockIdx.y;
  const int64_t* __restrict__ keys = a.keys[w];
  int32_t* __restrict__ cursor = a.cursor[w];
  int32_t* __restrict__ perm = a.perm[w];
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pairs; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = keys[e];
    if (k < 0 || k >= n_nodes) continue;
    perm[atomicAdd(&cursor[k], 1)] = (int32_t)e;
  }
}

__device__ __forceinline__ void sort_segment(int32_t* a, int d);
__global__ void csr_sort2_kernel(CsrPair a, int64_t n_nodes) {
  const int w = blockIdx.y;
  const int32_t* __restrict__ rowptr = a.rowptr[w];
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = rowptr[n], e = rowptr[n + 1];
    sort_segment(a.perm[w] + b, e - b);
  }
}

__device__ __forceinline__ void sort_segment(int32_t* a, int d) {
  if (d <= 1) return;
  if (d <= 24) {
    for (int i = 1; i < d; ++i) {
      int32_t v = a[i]; int j = i - 1;
      while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
      a[j + 1] = v;
    }
    return;
  }
  for (int start = d / 2 - 1; start >= 0; --start) {  // heapsort
    int root = start; int32_t v = a[root];
    while (true) {
      int child = 2 * root + 1; if (child >= d) break;
      if (child + 1 < d && a[child + 1] > a[child]) ++child;
      if (a[child] <= v) break;
      a[root] = a[child]; root = child;
    }
    a[root] = v;
  }
  for (int end = d - 1; end > 0; --end) {
    int32_t v = a[end]; a[end] = a[0];
    int root = 0;
    while (true) {
      int child = 2 * root + 1; if (child >= end) break;
      if (child + 1 < end && a[child + 1] > a[child]) ++child;
      if (a[child] <= v) break;
      a[root] = a[child]; root = child;
    }
    a[root] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
template <int V>
__global__ void gather_diff_kernel(const float* __restrict__ h, const int64_t* __restrict__ src,
                                   const int64_t* __restrict__ dst, float* __restrict__ out, int64_t n_pairs,
                                   int64_t width) {
  const int64_t vpr = width / V, total = n_pairs * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / vpr, v = idx - p * vpr;
    const float* hi = h + dst[p] * width + v * V;
    const float* hj = h + src[p] * width + v * V;
    float* o = out + p * width + v * V;
    if constexpr (V == 4) {
      float4 a = *reinterpret_cast<const float4*>(hi), b = *reinterpret_cast<const float4*>(hj);
      *reinterpret_cast<float4*>(o) = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    } else {
      o[0] = hi[0] - hj[0];
    }
  }
}

template <int V>
__global__ void segment_reduce_kernel(const float* __restrict__ msg, const int32_t* __restrict__ rowptr,
                                      const int32_t* __restrict__ perm, float* __restrict__ out, int64_t n_nodes,
                                      int64_t width, int mean) {
  const int64_t vpr = width / V, total = n_nodes * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / vpr, v = idx - n * vpr;
    const int32_t b = rowptr[n], e = rowptr[n + 1];
    float acc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = 0.f;
    for (int32_t p = b; p < e; ++p) {
      const float* m = msg + (int64_t)perm[p] * width + v * V;
      if constexpr (V == 4) {
        float4 t = *reinterpret_cast<const float4*>(m);
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      } else acc[0] += m[0];
    }
    if (mean) {
      const float d = (float)(e - b > 1 ? e - b : 1);
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] /= d;
    }
    float* o = out + n * width + v * V;
    if constexpr (V == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else o[0] = acc[0];
  }
}

template <int V>
__global__ void scatter_diff_kernel(const float* __restrict__ g, const int32_t* __restrict__ rp_dst,
                                    const int32_t* __restrict__ pm_dst, const int32_t* __restrict__ rp_src,
                                    const int32_t* __restrict__ pm_src, float* __restrict__ gh, int64_t n_nodes,
                                    int64_t width, int accumulate) {
  const int64_t vpr = width / V, total = n_nodes * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / vpr, v = idx - n * vpr;
    float acc[V];
    float* o = gh + n * width + v * V;
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = accumulate ? o[q] : 0.f;
    for (int32_t p = rp_dst[n]; p < rp_dst[n + 1]; ++p) {
      const float* m = g + (int64_t)pm_dst[p] * width + v * V;
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] += m[q];
    }
    for (int32_t p = rp_src[n]; p < rp_src[n + 1]; ++p) {
      const float* m = g + (int64_t)pm_src[p] * width + v * V;
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] -= m[q];
    }
#pragma unroll
    for (int q = 0; q < V; ++q) o[q] = acc[q];
  }
}

template <int V>
__global__ void segment_expand_kernel(const float* __restrict__ go, const int64_t* __restrict__ dst,
                                      const int32_t* __restrict__ rowptr, float* __restrict__ gm, int64_t n_pairs,
                                      int64_t width, int mean) {
  const int64_t vpr = width / V, total = n_pairs * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / vpr, v = idx - p * vpr;
    const int64_t n = dst[p];
    float sc = 1.f;
    if (mean) { int32_t d = rowptr[n + 1] - rowptr[n]; sc = 1.f / (float)(d > 1 ? d : 1); }
    const float* s = go + n * width + v * V;
    float* o = gm + p * width + v * V;
#pragma unroll
    for (int q = 0; q < V; ++q) o[q] = s[q] * sc;
  }
}


// ---------------------------------------------------------------------------------------------------
// receiver-sorted helpers for the fused EGCL path
__global__ void sorted_indices_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                      const int32_t* __restrict__ perm, int32_t* __restrict__ ss, int32_t* __restrict__ ds,
                                      int32_t* __restrict__ rank, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = perm[p];
    ss[p] = (int32_t)src[e];
    ds[p] = (int32_t)dst[e];
    if (rank) rank[e] = (int32_t)p;
  }
}

__global__ void rank_kernel(const int32_t* __restrict__ perm, int32_t* __restrict__ rank, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) rank[perm[p]] = (int32_t)p;
}

__global__ void segment_reduce_sorted_kernel(const float4* __restrict__ msg, const int32_t* __restrict__ rowptr,
                                             float4* __restrict__ out, int64_t n_nodes, int64_t vpr, int mean) {
  const int64_t total = n_nodes * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / vpr, v = idx - n * vpr;
    const int32_t b = rowptr[n], e = rowptr[n + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t p = b; p < e; ++p) {
      const float4 t = msg[(int64_t)p * vpr + v];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (mean) {
      const float d = (float)(e - b > 1 ? e - b : 1);
      acc.x /= d; acc.y /= d; acc.z /= d; acc.w /= d;
    }
    out[idx] = acc;
  }
}

__global__ void segment_expand_sorted_kernel(const float4* __restrict__ go, const int32_t* __restrict__ dst_sorted,
                                             const int32_t* __restrict__ rowptr, float4* __restrict__ gm, int64_t n_pairs,
                                             int64_t vpr, int mean) {
  const int64_t total = n_pairs * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / vpr, v = idx - p * vpr;
    const int32_t n = dst_sorted[p];
    float sc = 1.f;
    if (mean) { const int32_t d = rowptr[n + 1] - rowptr[n]; sc = 1.f / (float)(d > 1 ? d : 1); }
    float4 t = go[(int64_t)n * vpr + v];
    t.x *= sc; t.y *= sc; t.z *= sc; t.w *= sc;
    gm[idx] = t;
  }
}

__global__ void scatter_diff_sorted_kernel(const float* __restrict__ g, int64_t ld, const int32_t* __restrict__ rp_dst,
                                           const int32_t* __restrict__ rp_src, const int32_t* __restrict__ pm_src,
                                           const int32_t* __restrict__ rank, float* __restrict__ gh, int64_t n_nodes,
                                           int64_t width, int accumulate) {
  const int64_t vpr = width / 4, total = n_nodes * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / vpr, v = idx - n * vpr;
    float4* o = reinterpret_cast<float4*>(gh + n * width) + v;
    float4 acc = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t p = rp_dst[n]; p < rp_dst[n + 1]; ++p) {
      const float4 t = *(reinterpret_cast<const float4*>(g + (int64_t)p * ld) + v);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    for (int32_t q = rp_src[n]; q < rp_src[n + 1]; ++q) {
      const int64_t p = rank[pm_src[q]];
      const float4 t = *(reinterpret_cast<const float4*>(g + p * ld) + v);
      acc.x -= t.x; acc.y -= t.y; acc.z -= t.z; acc.w -= t.w;
    }
    *o = acc;
  }
}

// out[n, :] = sum_{pairs p with dst = n} g[p, col_dst : col_dst + width] + sum_{pairs with src = n} g[p, col_src : ...]
// (gradient of a per-simplex table that entered every pair as  table[src] | table[dst]); fixed order, no atomics
__global__ void scatter_pair_sorted_kernel(const float* __restrict__ g, int64_t ld, int64_t col_src, int64_t col_dst,
                                           const int32_t* __restrict__ rp_dst, const int32_t* __restrict__ rp_src,
                                           const int32_t* __restrict__ pm_src, const int32_t* __restrict__ rank,
                                           float* __restrict__ out, int64_t n_nodes, int64_t width) {
  const int64_t vpr = width / 4, total = n_nodes * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / vpr, v = idx - n * vpr;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t p = rp_dst[n]; p < rp_dst[n + 1]; ++p) {
      const float4 t = *(reinterpret_cast<const float4*>(g + (int64_t)p * ld + col_dst) + v);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    for (int32_t q = rp_src[n]; q < rp_src[n + 1]; ++q) {
      const int64_t p = rank[pm_src[q]];
      const float4 t = *(reinterpret_cast<const float4*>(g + p * ld + col_src) + v);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *(reinterpret_cast<float4*>(out + n * width) + v) = acc;
  }
}

// out[r, :] = a[r, :] + b[r, :] + c[r, :] with per-operand leading dimensions (b / c may be column slices of wider rows)
__global__ void add3_rows_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                 const float* __restrict__ c, int64_t ldc, float* __restrict__ out, int64_t n_rows, int64_t width) {
  const int64_t vpr = width / 4, total = n_rows * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / vpr, v = idx - r * vpr;
    const float4 x = *(reinterpret_cast<const float4*>(a + r * lda) + v);
    const float4 y = *(reinterpret_cast<const float4*>(b + r * ldb) + v);
    const float4 z = *(reinterpret_cast<const float4*>(c + r * ldc) + v);
    *(reinterpret_cast<float4*>(out + r * width) + v) = make_float4(x.x + y.x + z.x, x.y + y.y + z.y, x.z + y.z + z.z, x.w + y.w + z.w);
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ g, int64_t ld, int64_t col0, const int32_t* __restrict__ eid,
                                    float* __restrict__ out, int64_t n_rows, int64_t width) {
  const int64_t vpr = width / 4, total = n_rows * vpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / vpr, v = idx - p * vpr;
    const float4 t = *(reinterpret_cast<const float4*>(g + p * ld + col0) + v);
    *(reinterpret_cast<float4*>(out + (int64_t)eid[p] * width) + v) = t;
  }
}

inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count_cached() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace csmpn

using namespace csmpn;

extern "C" {

int64_t csmpn_csr_workspace(int64_t n_pairs, int64_t n_nodes) {
  (void)n_pairs;
  if (n_nodes < 0) return 0;
  int64_t tiles = (n_nodes + kScanTile - 1) / kScanTile + 1;
  // cursor[n_nodes] + tile_sums[tiles] + total + bad flag, 16B aligned chunks
  return ((n_nodes + 4) + (tiles + 4) + 8) * (int64_t)sizeof(int32_t);
}

int csmpn_csr_build(const int64_t* keys, int64_t n_pairs, int64_t n_nodes, int32_t* rowptr, int32_t* perm,
                    void* workspace, int64_t workspace_bytes, csmpn_stream_t stream) {
  if (n_pairs < 0 || n_nodes < 0 || !rowptr || (n_pairs > 0 && (!keys || !perm))) return CSMPN_ERR_BAD_ARG;
  if (n_pairs >= (int64_t)1 << 31 || n_nodes >= ((int64_t)1 << 31) - 1) return CSMPN_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < csmpn_csr_workspace(n_pairs, n_nodes)) return CSMPN_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* cursor = (int32_t*)workspace;
  const int64_t tiles = (n_nodes + kScanTile - 1) / kScanTile;
  int32_t* tile_sums = cursor + (n_nodes + 4) / 4 * 4;
  int32_t* total = tile_sums + (tiles + 4) / 4 * 4;
  int32_t* bad = total + 1;
  // counts live in cursor[] first
  CSMPN_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)csmpn_csr_workspace(n_pairs, n_nodes), s));
  if (n_pairs > 0) {
    csr_count_kernel<<<grid_for(n_pairs, 256), 256, 0, s>>>(keys, n_pairs, n_nodes, cursor, bad);
    CSMPN_LAUNCH_CHECK("csr_count");
  }
  if (n_nodes > 0) {
    scan_tiles_kernel<<<(unsigned)tiles, kScanThreads, 0, s>>>(cursor, rowptr, tile_sums, n_nodes);
    CSMPN_LAUNCH_CHECK("scan_tiles");
  }
  scan_sums_kernel<<<1, 1024, 0, s>>>(tile_sums, tiles, total);
  CSMPN_LAUNCH_CHECK("scan_sums");
  scan_apply_kernel<<<grid_for(n_nodes + 1, 256), 256, 0, s>>>(rowptr, tile_sums, n_nodes, cursor, total);
  CSMPN_LAUNCH_CHECK("scan_apply");
  if (n_pairs > 0) {
    csr_fill_kernel<<<grid_for(n_pairs, 256), 256, 0, s>>>(keys, n_pairs, n_nodes, cursor, perm);
    CSMPN_LAUNCH_CHECK("csr_fill");
    csr_sort_segments_kernel<<<grid_for(n_nodes, 128), 128, 0, s>>>(rowptr, n_nodes, perm);
    CSMPN_LAUNCH_CHECK("csr_sort_segments");
  }
  return CSMPN_OK;
}

#define CSMPN_VEC_DISPATCH(width, KERNEL, total, ...)                                     \
  if ((width) % 4 == 0) KERNEL<4><<<grid_for((total) / 4, 256), 256, 0, s>>>(__VA_ARGS__); \
  else KERNEL<1><<<grid_for((total), 256), 256, 0, s>>>(__VA_ARGS__);

int csmpn_gather_diff(const float* h, const int64_t* src, const int64_t* dst, float* out, int64_t n_pairs, int64_t width,
                      csmpn_stream_t stream) {
  if (n_pairs < 0 || width <= 0) return CSMPN_ERR_BAD_ARG;
  if (n_pairs == 0) return CSMPN_OK;
  if (!h || !src || !dst || !out) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_VEC_DISPATCH(width, gather_diff_kernel, n_pairs * width, h, src, dst, out, n_pairs, width);
  CSMPN_LAUNCH_CHECK("gather_diff");
  return CSMPN_OK;
}

int csmpn_segment_reduce(const float* msg, const int32_t* rowptr, const int32_t* perm, float* out, int64_t n_nodes,
                         int64_t width, int mean, csmpn_stream_t stream) {
  if (n_nodes < 0 || width <= 0) return CSMPN_ERR_BAD_ARG;
  if (n_nodes == 0) return CSMPN_OK;
  if (!rowptr || !out) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_VEC_DISPATCH(width, segment_reduce_kernel, n_nodes * width, msg, rowptr, perm, out, n_nodes, width, mean);
  CSMPN_LAUNCH_CHECK("segment_reduce");
  return CSMPN_OK;
}

int csmpn_scatter_diff(const float* g, const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* rowptr_src,
                       const int32_t* perm_src, float* grad_h, int64_t n_nodes, int64_t width, int accumulate,
                       csmpn_stream_t stream) {
  if (n_nodes < 0 || width <= 0) return CSMPN_ERR_BAD_ARG;
  if (n_nodes == 0) return CSMPN_OK;
  if (!rowptr_dst || !rowptr_src || !grad_h) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_VEC_DISPATCH(width, scatter_diff_kernel, n_nodes * width, g, rowptr_dst, perm_dst, rowptr_src, perm_src, grad_h,
                     n_nodes, width, accumulate);
  CSMPN_LAUNCH_CHECK("scatter_diff");
  return CSMPN_OK;
}

int csmpn_segment_expand(const float* grad_out, const int64_t* dst, const int32_t* rowptr, float* grad_msg,
                         int64_t n_pairs, int64_t width, int mean, csmpn_stream_t stream) {
  if (n_pairs < 0 || width <= 0) return CSMPN_ERR_BAD_ARG;
  if (n_pairs == 0) return CSMPN_OK;
  if (!grad_out || !dst || !rowptr || !grad_msg) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_VEC_DISPATCH(width, segment_expand_kernel, n_pairs * width, grad_out, dst, rowptr, grad_msg, n_pairs, width, mean);
  CSMPN_LAUNCH_CHECK("segment_expand");
  return CSMPN_OK;
}

int csmpn_csr_build_pair(const int64_t* src, const int64_t* dst, int64_t n_pairs, int64_t n_nodes, int32_t* rowptr_dst,
                         int32_t* perm_dst, int32_t* rowptr_src, int32_t* perm_src, int32_t* src_sorted, int32_t* dst_sorted,
                         int32_t* rank, void* workspace, int64_t workspace_bytes, csmpn_stream_t stream) {
  if (n_pairs < 0 || n_nodes < 0 || !rowptr_dst || !rowptr_src) return CSMPN_ERR_BAD_ARG;
  if (n_pairs > 0 && (!src || !dst || !perm_dst || !perm_src)) return CSMPN_ERR_BAD_ARG;
  if ((src_sorted == nullptr) != (dst_sorted == nullptr)) return CSMPN_ERR_BAD_ARG;
  if (n_nodes > (int64_t)kScan1Threads * kScan1Per || n_pairs >= ((int64_t)1 << 31)) return CSMPN_ERR_UNSUPPORTED;  // -> csmpn_csr_build
  const int64_t half = (n_nodes + 4) / 4 * 4;  // int32 counters per key array
  if (!workspace || workspace_bytes < 2 * half * (int64_t)sizeof(int32_t)) return CSMPN_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  CsrPair a;
  a.keys[0] = dst; a.keys[1] = src;
  a.rowptr[0] = rowptr_dst; a.rowptr[1] = rowptr_src;
  a.perm[0] = perm_dst; a.perm[1] = perm_src;
  a.cursor[0] = (int32_t*)workspace; a.cursor[1] = (int32_t*)workspace + half;
  CSMPN_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)(2 * half) * sizeof(int32_t), s));
  if (n_pairs > 0) {
    csr_count2_kernel<<<dim3(grid_for(n_pairs, 256), 2), 256, 0, s>>>(a, n_pairs, n_nodes);
    CSMPN_LAUNCH_CHECK("csr_count2");
  }
  csr_scan2_kernel<<<2, kScan1Threads, 0, s>>>(a, (int)n_nodes);
  CSMPN_LAUNCH_CHECK("csr_scan2");
  if (n_pairs > 0) {
    csr_fill2_kernel<<<dim3(grid_for(n_pairs, 256), 2), 256, 0, s>>>(a, n_pairs, n_nodes);
    CSMPN_LAUNCH_CHECK("csr_fill2");
    csr_sort2_kernel<<<dim3(grid_for(n_nodes, 128), 2), 128, 0, s>>>(a, n_nodes);
    CSMPN_LAUNCH_CHECK("csr_sort2");
    if (src_sorted) {
      sorted_indices_kernel<<<grid_for(n_pairs, 256), 256, 0, s>>>(src, dst, perm_dst, src_sorted, dst_sorted, rank, n_pairs);
      CSMPN_LAUNCH_CHECK("csr_sorted_indices");
    }
  }
  return CSMPN_OK;
}

int csmpn_csr_sorted_indices(const int64_t* src, const int64_t* dst, const int32_t* perm, int32_t* src_sorted,
                             int32_t* dst_sorted, int64_t n_pairs, csmpn_stream_t stream) {
  if (n_pairs < 0) return CSMPN_ERR_BAD_ARG;
  if (n_pairs == 0) return CSMPN_OK;
  if (!src || !dst || !perm || !src_sorted || !dst_sorted) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  sorted_indices_kernel<<<grid_for(n_pairs, 256), 256, 0, s>>>(src, dst, perm, src_sorted, dst_sorted, nullptr, n_pairs);
  CSMPN_LAUNCH_CHECK("csr_sorted_indices");
  return CSMPN_OK;
}

int csmpn_segment_reduce_sorted(const float* msg, const int32_t* rowptr, float* out, int64_t n_nodes, int64_t width,
                                int mean, csmpn_stream_t stream) {
  if (n_nodes < 0 || width <= 0 || width % 4) return CSMPN_ERR_BAD_ARG;
  if (n_nodes == 0) return CSMPN_OK;
  if (!rowptr || !out) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  segment_reduce_sorted_kernel<<<grid_for(n_nodes * width / 4, 256), 256, 0, s>>>(
      reinterpret_cast<const float4*>(msg), rowptr, reinterpret_cast<float4*>(out), n_nodes, width / 4, mean);
  CSMPN_LAUNCH_CHECK("segment_reduce_sorted");
  return CSMPN_OK;
}

int csmpn_segment_expand_sorted(const float* grad_out, const int32_t* dst_sorted, const int32_t* rowptr, float* grad_msg,
                                int64_t n_pairs, int64_t width, int mean, csmpn_stream_t stream) {
  if (n_pairs < 0 || width <= 0 || width % 4) return CSMPN_ERR_BAD_ARG;
  if (n_pairs == 0) return CSMPN_OK;
  if (!grad_out || !dst_sorted || !rowptr || !grad_msg) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  segment_expand_sorted_kernel<<<grid_for(n_pairs * width / 4, 256), 256, 0, s>>>(
      reinterpret_cast<const float4*>(grad_out), dst_sorted, rowptr, reinterpret_cast<float4*>(grad_msg), n_pairs,
      width / 4, mean);
  CSMPN_LAUNCH_CHECK("segment_expand_sorted");
  return CSMPN_OK;
}

int csmpn_scatter_diff_sorted(const float* g, int64_t ld, const int32_t* rowptr_dst, const int32_t* rowptr_src,
                              const int32_t* perm_src, const int32_t* rank, float* grad_h, int64_t n_nodes,
                              int64_t width, int accumulate, csmpn_stream_t stream) {
  if (n_nodes < 0 || width <= 0 || width % 4 || ld % 4 || ld < width) return CSMPN_ERR_BAD_ARG;
  if (n_nodes == 0) return CSMPN_OK;
  if (!rowptr_dst || !rowptr_src || !grad_h) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  scatter_diff_sorted_kernel<<<grid_for(n_nodes * width / 4, 256), 256, 0, s>>>(g, ld, rowptr_dst, rowptr_src, perm_src,
                                                                                  rank, grad_h, n_nodes, width, accumulate);
  CSMPN_LAUNCH_CHECK("scatter_diff_sorted");
  return CSMPN_OK;
}

int csmpn_scatter_rows(const float* g, int64_t ld, int64_t col0, const int32_t* eid, float* out, int64_t n_rows,
                       int64_t width, csmpn_stream_t stream) {
  if (n_rows < 0 || width <= 0 || width % 4 || ld % 4 || col0 % 4) return CSMPN_ERR_BAD_ARG;
  if (n_rows == 0) return CSMPN_OK;
  if (!g || !eid || !out) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  scatter_rows_kernel<<<grid_for(n_rows * width / 4, 256), 256, 0, s>>>(g, ld, col0, eid, out, n_rows, width);
  CSMPN_LAUNCH_CHECK("scatter_rows");
  return CSMPN_OK;
}

int csmpn_scatter_pair_sorted(const float* g, int64_t ld, int64_t col_src, int64_t col_dst, const int32_t* rowptr_dst,
                              const int32_t* rowptr_src, const int32_t* perm_src, const int32_t* rank, float* out,
                              int64_t n_nodes, int64_t width, csmpn_stream_t stream) {
  if (n_nodes < 0 || width <= 0 || width % 4 || ld % 4 || col_src % 4 || col_dst % 4 || col_src < 0 || col_dst < 0)
    return CSMPN_ERR_BAD_ARG;
  if (n_nodes == 0) return CSMPN_OK;
  if (!g || !rowptr_dst || !rowptr_src || !out) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  scatter_pair_sorted_kernel<<<grid_for(n_nodes * width / 4, 256), 256, 0, s>>>(g, ld, col_src, col_dst, rowptr_dst, rowptr_src,
                                                                                  perm_src, rank, out, n_nodes, width);
  CSMPN_LAUNCH_CHECK("scatter_pair_sorted");
  return CSMPN_OK;
}

int csmpn_add3_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* out,
                    int64_t n_rows, int64_t width, csmpn_stream_t stream) {
  if (n_rows < 0 || width <= 0 || width % 4 || lda % 4 || ldb % 4 || ldc % 4) return CSMPN_ERR_BAD_ARG;
  if (n_rows == 0) return CSMPN_OK;
  if (!a || !b || !c || !out) return CSMPN_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  add3_rows_kernel<<<grid_for(n_rows * width / 4, 256), 256, 0, s>>>(a, lda, b, ldb, c, ldc, out, n_rows, width);
  CSMPN_LAUNCH_CHECK("add3_rows");
  return CSMPN_OK;
}

int csmpn_csr_rank(const int32_t* perm, int32_t* rank, int64_t n_pairs, csmpn_stream_t stream) {
  if (n_pairs < 0) return CSMPN_ERR_BAD_ARG;
  if (n_pairs == 0) return CSMPN_OK;
  if (!perm || !rank) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  rank_kernel<<<grid_for(n_pairs, 256), 256, 0, s>>>(perm, rank, n_pairs);
  CSMPN_LAUNCH_CHECK("csr_rank");
  return CSMPN_OK;
}

}  // extern "C"
