// Simplicial lifting on the GPU (north-star item 4): vertices -> edges -> triangles, the per-dimension simplex
// numbering, x_ind / node_types and the six adjacency blocks merged into ONE edge_index, bit-exact (values AND order)
// to the reference's CPU pipeline
//   csmpn/data/modules/utils.py:63-136 (rips_lift + generate_*_single), :151-207,250-388 (simplicial_lift +
//   generate_*), :210-248 (simplicial_lift_hulls) and csmpn/data/modules/simplicial_data.py:105-175,218-222
//   (add_missing_adj, get_edge, x_ind, node_types), :254-302 (ManualTransform).
//
// The reference walks a gudhi SimplexTree in Python.  For complexes of dimension <= 2 everything it produces has a
// closed form over the vertex adjacency bit masks (n <= 32 vertices per complex, one 32-bit mask per vertex):
//   numbering    per dimension, lexicographic over sorted vertex tuples (depth-first tree traversal)
//   x_ind row    the vertices of the simplex in CPython frozenset iteration order (8-slot open-addressing table,
//                slot = v & 7, collision -> i = (5 i + 1 + perturb) & 7 with perturb >>= 5), as float32, zero padded
//   0_0          for v ascending, for every neighbour u ascending: (u, v)                       [upper adjacency]
//                + (rips / hulls only, utils.py:90-96) for i, for j: every ordered pair i != j that is not a sorted
//                  edge [i, j] -- all non-edges both ways plus (big, small) of every edge a second time
//   0_1 / 1_0    for edge e = (a, b) ascending: (a, e), (b, e)         / the same columns reversed
//   1_1          for e ascending, for every triangle t = e + {x}, x ascending, for the faces of t in the order
//                (p,q), (p,r), (q,r) (gudhi boundary order: drop the largest vertex first) except e itself: (face, e)
//   1_2 / 2_1    for t = (p,q,r) ascending: ((p,q), t), ((p,r), t), ((q,r), t)   / reversed
// with ids offset by the per-dimension simplex counts and blocks concatenated in the order 0_0,0_1,1_0,1_1,1_2,2_1.
// One warp lifts one complex; a batch of complexes is lifted by one launch and comes out already collated
// (node offsets added, edge_index concatenated along the pair axis) the way PyG's DataLoader would collate them.
#include "common.cuh"

namespace csmpn {

constexpr int kLiftWarps = 4;      // complexes per CTA
constexpr int kMaxV = 32;          // vertices per complex of the one-word kernels (one 32-bit adjacency mask per vertex)
constexpr int kMaxV2 = 64;         // ... of the two-word kernels (uint64 masks), selected by csmpn_lift_desc.max_vertices > 32

constexpr int kMotionV = 31, kMotionE = 12, kMotionT = 4, kMotionFixedPairs = 96;
__constant__ uint8_t c_motion_edges[kMotionE][2] = {{6, 7}, {7, 8}, {6, 8}, {1, 2}, {2, 3}, {1, 3}, {24, 25}, {25, 26},
                                                    {24, 26}, {22, 23}, {21, 22}, {21, 23}};
__constant__ uint8_t c_motion_tris[kMotionT][3] = {{6, 7, 8}, {1, 2, 3}, {24, 25, 26}, {21, 22, 23}};

// M = mask word (uint32_t: up to 32 vertices, uint64_t: up to 64), KV = vertex capacity
template <typename M, int KV>
struct LiftWarpT {
  static constexpr int KE = KV * (KV - 1) / 2;
  M adj[KV];                  // neighbour mask of every vertex
  M adj0[KV];                 // clique modes with filters: the unfiltered graph (triangle candidates are ITS 3-cliques)
  M cof[KE];                  // per edge (a,b): every x such that {a,b,x} is a triangle of the complex
  uint16_t ebase[KV + 1];     // number of edges whose smaller vertex is < a
  uint32_t tbase[KE + 1];     // number of triangles whose smallest edge (p,q) precedes edge e
  uint32_t cbase[KE + 1];     // number of (triangle, edge) incidences of the edges preceding e
  uint8_t ea[KE], eb[KE];
  int n, n_e, n_t;
};

template <typename M> __device__ __forceinline__ M bit_of(int v) { return (M)1 << v; }
template <typename M> __device__ __forceinline__ M bits_above(int v) {  // {v+1 .. capacity-1}
  return v >= (int)(8 * sizeof(M)) - 1 ? (M)0 : ~((((M)2) << v) - (M)1);
}
template <typename M> __device__ __forceinline__ M bits_below(int v) { return (((M)1) << v) - (M)1; }  // {0 .. v-1}
__device__ __forceinline__ int popc(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int popc(uint64_t m) { return __popcll(m); }
__device__ __forceinline__ int ffs0(uint32_t m) { return __ffs(m) - 1; }
__device__ __forceinline__ int ffs0(uint64_t m) { return __ffsll((long long)m) - 1; }
__device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
__device__ __forceinline__ void atomic_or(uint64_t* p, uint64_t v) { atomicOr(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }

__device__ __forceinline__ bool clique_like(int mode) { return mode == CSMPN_LIFT_CLIQUE || mode == CSMPN_LIFT_KNN; }

// utils.py:139-148 triangle_area: 0.5 |(v2 - v1) x (v3 - v1)| in fp32 (2-D points are embedded with z = 0).  The reference
// calls torch.cross without `dim`, which for exactly three triangles picks the wrong axis; this is the per-triangle formula.
__device__ __forceinline__ float lift_tri_area(const float* P, int D, int a, int b, int c) {
  float u[3] = {0.f, 0.f, 0.f}, w[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < D && k < 3; ++k) {
    u[k] = __fsub_rn(P[(int64_t)b * D + k], P[(int64_t)a * D + k]);
    w[k] = __fsub_rn(P[(int64_t)c * D + k], P[(int64_t)a * D + k]);
  }
  const float cx = __fsub_rn(__fmul_rn(u[1], w[2]), __fmul_rn(u[2], w[1]));
  const float cy = __fsub_rn(__fmul_rn(u[2], w[0]), __fmul_rn(u[0], w[2]));
  const float cz = __fsub_rn(__fmul_rn(u[0], w[1]), __fmul_rn(u[1], w[0]));
  return 0.5f * sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
}
__device__ __forceinline__ float lift_tri_area_sorted(const float* P, int D, int a, int b, int c) {
  if (a > b) { const int t = a; a = b; b = t; }
  if (b > c) { const int t = b; b = c; c = t; }
  if (a > b) { const int t = a; a = b; b = t; }
  return lift_tri_area(P, D, a, b, c);
}
__device__ __forceinline__ float lift_edge_len(const float* P, int D, int a, int b) {
  float acc = 0.f;
  for (int k = 0; k < D; ++k) {
    const float df = __fsub_rn(P[(int64_t)a * D + k], P[(int64_t)b * D + k]);
    acc = __fadd_rn(acc, __fmul_rn(df, df));
  }
  return sqrtf(acc);
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// CPython set iteration order of up to three small non-negative ints inserted in ascending order.
__device__ __forceinline__ void frozenset_order(const int* v, int k, float* out) {
  uint64_t table = 0;  // 8 slots x 8 bits, value + 1 (0 = empty)
  for (int q = 0; q < k; ++q) {
    uint32_t perturb = (uint32_t)v[q], i = perturb & 7u;
    while ((table >> (8 * i)) & 0xffu) {
      perturb >>= 5;
      i = (i * 5u + 1u + perturb) & 7u;
    }
    table |= (uint64_t)(v[q] + 1) << (8 * i);
  }
  int m = 0;
  for (int i = 0; i < 8; ++i) {
    const int e = (int)((table >> (8 * i)) & 0xffu);
    if (e) out[m++] = (float)(e - 1);
  }
  for (; m < 3; ++m) out[m] = 0.f;
}

// index of the edge (u, w), u < w, in lexicographic order
template <typename M, int KV>
__device__ __forceinline__ int edge_index_of(const LiftWarpT<M, KV>& s, int u, int w) {
  return s.ebase[u] + popc((M)(s.adj[u] & bits_above<M>(u) & bits_below<M>(w)));
}

// Build the adjacency masks, the lexicographic edge list and the triangle tables of complex `c` (one warp; vertex v is
// handled by lane v % 32).
template <typename M, int KV>
__device__ void lift_build(LiftWarpT<M, KV>& s, const csmpn_lift_desc& d, int c, int lane) {
  const int v0 = d.vptr[c], n = d.vptr[c + 1] - v0;
  if (lane == 0) s.n = n;
  if (d.mode == CSMPN_LIFT_RIPS) {
    // gudhi RipsComplex: Euclidean distance in double over the float32 coordinates, edge iff dist <= max_edge_length
    for (int v = lane; v < KV; v += 32) {
      M mine = 0;
      if (v < n) {
        for (int j = 0; j < n; ++j) {
          if (j == v) continue;
          double acc = 0.0;
          for (int k = 0; k < d.point_dim; ++k) {
            const double diff = __dsub_rn((double)d.points[(int64_t)(v0 + v) * d.point_dim + k],
                                          (double)d.points[(int64_t)(v0 + j) * d.point_dim + k]);
            acc = __dadd_rn(acc, __dmul_rn(diff, diff));
          }
          if (sqrt(acc) <= d.max_edge_length) mine |= bit_of<M>(j);
        }
      }
      s.adj[v] = mine;
    }
  } else if (clique_like(d.mode)) {
    for (int v = lane; v < KV; v += 32) s.adj[v] = 0;
    __syncwarp();
    if (d.mode == CSMPN_LIFT_CLIQUE) {
      const int64_t p0 = d.pptr[c], p1 = d.pptr[c + 1];
      for (int64_t p = p0 + lane; p < p1; p += 32) {
        const int a = (int)d.pairs[p], b = (int)d.pairs[d.n_pairs + p];
        if (a != b && a >= 0 && b >= 0 && a < n && b < n) {
          atomic_or(&s.adj[a], bit_of<M>(b));
          atomic_or(&s.adj[b], bit_of<M>(a));
        }
      }
    } else {
      // knn_graph(points, k) (csmpn/data/md17.py:64, nba.py:48; torch_cluster): every vertex is joined to its k nearest
      // other vertices, nearest first, ties to the smaller index; the lift only uses the UNDIRECTED edge set
      // (nx.Graph, utils.py:171-172).  Distances in double over the fp32 coordinates.
      for (int v = lane; v < n; v += 32) {
        M taken = bit_of<M>(v);
        const int kk = d.knn_k < n - 1 ? d.knn_k : n - 1;
        for (int t = 0; t < kk; ++t) {
          double best = 0.0;
          int bj = -1;
          for (int j = 0; j < n; ++j) {
            if ((taken >> j) & 1) continue;
            double acc = 0.0;
            for (int k = 0; k < d.point_dim; ++k) {
              const double diff = __dsub_rn((double)d.points[(int64_t)(v0 + v) * d.point_dim + k],
                                            (double)d.points[(int64_t)(v0 + j) * d.point_dim + k]);
              acc = __dadd_rn(acc, __dmul_rn(diff, diff));
            }
            const double dist = sqrt(acc);
            if (bj < 0 || dist < best) { best = dist; bj = j; }
          }
          taken |= bit_of<M>(bj);
          atomic_or(&s.adj[v], bit_of<M>(bj));
          atomic_or(&s.adj[bj], bit_of<M>(v));
        }
      }
    }
    __syncwarp();
    if (d.use_filters) {
      // utils.py:181-200: an edge of the graph enters the complex iff its length is <= edge_th OR it is a face of a kept
      // triangle (SimplexTree.insert adds all faces); a 3-clique of the graph is kept iff its area is <= tri_th
      const float* P = d.points + (int64_t)v0 * d.point_dim;
      for (int v = lane; v < KV; v += 32) s.adj0[v] = v < n ? s.adj[v] : (M)0;
      __syncwarp();
      M keepv[KV / 32];
#pragma unroll
      for (int i = 0; i < KV / 32; ++i) {
        const int v = lane + 32 * i;
        M keep = 0;
        M m = v < n ? s.adj0[v] : (M)0;
        while (m) {
          const int b = ffs0(m);
          m &= m - 1;
          bool k = lift_edge_len(P, d.point_dim, v, b) <= d.edge_th;
          if (!k && d.max_dim >= 2) {
            M t = s.adj0[v] & s.adj0[b];
            while (t && !k) {
              const int x = ffs0(t);
              t &= t - 1;
              k = lift_tri_area_sorted(P, d.point_dim, v, b, x) <= d.tri_th;
            }
          }
          if (k) keep |= bit_of<M>(b);
        }
        keepv[i] = keep;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < KV / 32; ++i) s.adj[lane + 32 * i] = keepv[i];
    }
  } else {  // CSMPN_LIFT_FACETS: two vertices are joined iff some facet holds both
    for (int v = lane; v < KV; v += 32) {
      M mine = 0;
      if (v < n) {
        for (int f = d.fptr[c]; f < d.fptr[c + 1]; ++f) {
          M m = 0;
          for (int k = 0; k < d.facet_size; ++k) m |= bit_of<M>((int)d.facets[(int64_t)f * d.facet_size + k]);
          if ((m >> v) & 1) mine |= m;
        }
        mine &= ~bit_of<M>(v);
      }
      s.adj[v] = mine;
    }
  }
  __syncwarp();
  // edges, lexicographic: vertex a owns the edges (a, b > a); vertices in rounds of 32 (ascending)
  int ecarry = 0;
#pragma unroll
  for (int i = 0; i < KV / 32; ++i) {
    const int v = lane + 32 * i;
    const M up = v < n ? (M)(s.adj[v] & bits_above<M>(v)) : (M)0;
    int tot;
    const int eb0 = ecarry + warp_excl_scan(popc(up), lane, &tot);
    s.ebase[v] = (uint16_t)eb0;
    M m = up;
    int e = eb0;
    while (m) {
      const int b = ffs0(m);
      m &= m - 1;
      s.ea[e] = (uint8_t)v;
      s.eb[e] = (uint8_t)b;
      ++e;
    }
    ecarry += tot;
  }
  const int n_e = ecarry;
  if (lane == 0) { s.ebase[KV] = (uint16_t)n_e; s.n_e = n_e; }
  __syncwarp();
  // triangles: clique complexes close every 3-clique; the facet mode only keeps triples inside one facet
  int tcarry = 0, ccarry = 0;
  for (int e0 = 0; e0 < n_e; e0 += 32) {
    const int e = e0 + lane;
    M cand = 0;
    int a = 0, b = 0;
    if (e < n_e) {
      a = s.ea[e];
      b = s.eb[e];
      cand = s.adj[a] & s.adj[b];
      if (clique_like(d.mode) && d.use_filters) {  // 3-cliques of the UNFILTERED graph whose area passes tri_th
        const float* P = d.points + (int64_t)d.vptr[c] * d.point_dim;
        M t = s.adj0[a] & s.adj0[b], keep = 0;
        while (t) {
          const int x = ffs0(t);
          t &= t - 1;
          if (lift_tri_area_sorted(P, d.point_dim, a, b, x) <= d.tri_th) keep |= bit_of<M>(x);
        }
        cand = keep;
      }
      if (d.mode == CSMPN_LIFT_FACETS) {
        M keep = 0;
        for (int f = d.fptr[c]; f < d.fptr[c + 1]; ++f) {
          M m = 0;
          for (int k = 0; k < d.facet_size; ++k) m |= bit_of<M>((int)d.facets[(int64_t)f * d.facet_size + k]);
          if (((m >> a) & 1) && ((m >> b) & 1)) keep |= m;
        }
        cand &= keep;
      }
      if (d.max_dim < 2) cand = 0;
      s.cof[e] = cand;
    }
    int tt, ct;
    const int tb = warp_excl_scan(popc((M)(cand & bits_above<M>(b))), lane, &tt);
    const int cb = warp_excl_scan(popc(cand), lane, &ct);
    if (e < n_e) {
      s.tbase[e] = (uint32_t)(tcarry + tb);
      s.cbase[e] = (uint32_t)(ccarry + cb);
    }
    tcarry += tt;
    ccarry += ct;
  }
  if (lane == 0) {
    s.tbase[n_e] = (uint32_t)tcarry;
    s.cbase[n_e] = (uint32_t)ccarry;
    s.n_t = tcarry;
  }
  __syncwarp();
}

__device__ __forceinline__ int64_t pairs_of(int mode, int n, int n_e, int n_t) {
  int64_t p = 6ll * n_e + 12ll * n_t;
  if (!clique_like(mode)) p += (int64_t)n * (n - 1) - n_e;  // the extra 0_0 pairs of utils.py:90-96
  return p;
}

template <typename M, int KV, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) lift_count_kernel(csmpn_lift_desc d, int32_t* __restrict__ counts,
                                                                 int32_t* __restrict__ bad) {
  __shared__ LiftWarpT<M, KV> sm[WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * WARPS + w;
  if (c >= d.n_complexes) return;
  if (d.mode == CSMPN_LIFT_MOTION) {
    if (lane == 0) {
      counts[2 * c] = kMotionE;
      counts[2 * c + 1] = kMotionT;
      if (d.vptr[c + 1] - d.vptr[c] != kMotionV) *bad = 1;
    }
    return;
  }
  const int n = d.vptr[c + 1] - d.vptr[c];
  if (n > KV || n < 0) {
    if (lane == 0) { *bad = 1; counts[2 * c] = 0; counts[2 * c + 1] = 0; }
    return;
  }
  lift_build<M, KV>(sm[w], d, c, lane);
  if (lane == 0) {
    counts[2 * c] = sm[w].n_e;
    counts[2 * c + 1] = sm[w].n_t;
  }
}

// node_ptr[c] / pair_ptr[c]: first simplex / first pair of complex c in the collated batch (single CTA, chunked)
__global__ void __launch_bounds__(1024) lift_scan_kernel(csmpn_lift_desc d, const int32_t* __restrict__ counts,
                                                         int64_t* __restrict__ node_ptr, int64_t* __restrict__ pair_ptr) {
  __shared__ long long wsum_n[32], wsum_p[32];
  __shared__ long long carry_n, carry_p;
  if (threadIdx.x == 0) { carry_n = 0; carry_p = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < d.n_complexes; base += blockDim.x) {
    const int c = base + threadIdx.x;
    long long vn = 0, vp = 0;
    if (c < d.n_complexes) {
      const int n = d.vptr[c + 1] - d.vptr[c], n_e = counts[2 * c], n_t = counts[2 * c + 1];
      vn = n + n_e + n_t;
      vp = d.mode == CSMPN_LIFT_MOTION ? (long long)(d.pptr[c + 1] - d.pptr[c]) + kMotionFixedPairs : pairs_of(d.mode, n, n_e, n_t);
    }
    long long in = vn, ip = vp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long tn = __shfl_up_sync(0xffffffffu, in, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
      if (lane >= o) { in += tn; ip += tp; }
    }
    if (lane == 31) { wsum_n[wid] = in; wsum_p[wid] = ip; }
    __syncthreads();
    if (wid == 0) {
      long long a = wsum_n[lane], b = wsum_p[lane], ai = a, bi = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long ta = __shfl_up_sync(0xffffffffu, ai, o), tb = __shfl_up_sync(0xffffffffu, bi, o);
        if (lane >= o) { ai += ta; bi += tb; }
      }
      wsum_n[lane] = ai - a;
      wsum_p[lane] = bi - b;
    }
    __syncthreads();
    const long long en = carry_n + wsum_n[wid] + in - vn, ep = carry_p + wsum_p[wid] + ip - vp;
    if (c < d.n_complexes) { node_ptr[c] = en; pair_ptr[c] = ep; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) { carry_n = en + vn; carry_p = ep + vp; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { node_ptr[d.n_complexes] = carry_n; pair_ptr[d.n_complexes] = carry_p; }
}

struct LiftOut {
  int64_t* src;   // edge_index[0]
  int64_t* dst;   // edge_index[1]
  float* x_ind;   // [N, 3]
  int64_t* node_types;
  int64_t* batch;
};

__device__ __forceinline__ void put_pair(const LiftOut& o, int64_t pos, int64_t a, int64_t b) {
  o.src[pos] = a;
  o.dst[pos] = b;
}

template <typename M, int KV, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) lift_fill_kernel(csmpn_lift_desc d, const int64_t* __restrict__ node_ptr,
                                                                const int64_t* __restrict__ pair_ptr, LiftOut o) {
  __shared__ LiftWarpT<M, KV> sm[WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * WARPS + w;
  if (c >= d.n_complexes) return;
  const int64_t nb = node_ptr[c];
  int64_t pp = pair_ptr[c];
  if (d.mode == CSMPN_LIFT_MOTION) {
    // simplicial_data.py:263-298: skeleton pairs, then 2->1 | 1->2, then 1->0 | 0->1, then 1<->1 inside each triangle
    const int64_t p0 = d.pptr[c], nbp = d.pptr[c + 1] - p0;
    for (int64_t p = lane; p < nbp; p += 32) put_pair(o, pp + p, nb + d.pairs[p0 + p], nb + d.pairs[d.n_pairs + p0 + p]);
    pp += nbp;
    const int E0 = kMotionV, T0 = kMotionV + kMotionE;
    if (lane < 12) {
      put_pair(o, pp + lane, nb + T0 + lane / 3, nb + E0 + lane);
      put_pair(o, pp + 12 + lane, nb + E0 + lane, nb + T0 + lane / 3);
    }
    if (lane < 24) {
      const int e = lane >> 1, v = c_motion_edges[e][lane & 1];
      put_pair(o, pp + 24 + lane, nb + E0 + e, nb + v);
      put_pair(o, pp + 48 + lane, nb + v, nb + E0 + e);
    }
    if (lane < 24) {  // per triangle: ordered pairs (x, y), x != y, x-major
      const int t = lane / 6, q = lane % 6, x = q >> 1, yy = q & 1, y = yy + (yy >= x ? 1 : 0);
      put_pair(o, pp + 72 + lane, nb + E0 + 3 * t + x, nb + E0 + 3 * t + y);
    }
    for (int i = lane; i < kMotionV + kMotionE + kMotionT; i += 32) {
      float r[3] = {0.f, 0.f, 0.f};
      int ty = 0;
      if (i < kMotionV) r[0] = (float)i;
      else if (i < T0) { ty = 1; r[0] = c_motion_edges[i - E0][0]; r[1] = c_motion_edges[i - E0][1]; }
      else { ty = 2; r[0] = c_motion_tris[i - T0][0]; r[1] = c_motion_tris[i - T0][1]; r[2] = c_motion_tris[i - T0][2]; }
      o.x_ind[(nb + i) * 3] = r[0]; o.x_ind[(nb + i) * 3 + 1] = r[1]; o.x_ind[(nb + i) * 3 + 2] = r[2];
      o.node_types[nb + i] = ty;
      o.batch[nb + i] = c;
    }
    return;
  }
  LiftWarpT<M, KV>& s = sm[w];
  lift_build<M, KV>(s, d, c, lane);
  const int n = s.n, n_e = s.n_e, n_t = s.n_t;
  const int64_t E0 = nb + n, T0 = nb + n + n_e;
  // ---- simplices: x_ind, node_types, batch
  for (int v = lane; v < n; v += 32) {
    o.x_ind[(nb + v) * 3] = (float)v; o.x_ind[(nb + v) * 3 + 1] = 0.f; o.x_ind[(nb + v) * 3 + 2] = 0.f;
    o.node_types[nb + v] = 0;
    o.batch[nb + v] = c;
  }
  for (int e = lane; e < n_e; e += 32) {
    int v[2] = {s.ea[e], s.eb[e]};
    float r[3];
    frozenset_order(v, 2, r);
    o.x_ind[(E0 + e) * 3] = r[0]; o.x_ind[(E0 + e) * 3 + 1] = r[1]; o.x_ind[(E0 + e) * 3 + 2] = r[2];
    o.node_types[E0 + e] = 1;
    o.batch[E0 + e] = c;
  }
  // ---- block 0_0: upper adjacency through the shared edge (receivers in ascending order, rounds of 32 vertices) ...
  {
    int carry = 0;
#pragma unroll
    for (int i = 0; i < KV / 32; ++i) {
      const int v = lane + 32 * i;
      const M m0 = v < n ? s.adj[v] : (M)0;
      int tot;
      int pos = carry + warp_excl_scan(popc(m0), lane, &tot);
      M m = m0;
      while (m) {
        const int u = ffs0(m);
        m &= m - 1;
        put_pair(o, pp + pos++, nb + u, nb + v);
      }
      carry += tot;
    }
    pp += carry;  // = 2 n_e
    if (!clique_like(d.mode)) {  // ... plus the extra pairs of generate_adjacencies_single (utils.py:90-96)
      int carry2 = 0;
#pragma unroll
      for (int i = 0; i < KV / 32; ++i) {
        const int v = lane + 32 * i;
        const M up = v < n ? (M)(s.adj[v] & bits_above<M>(v)) : (M)0;
        int tot2;
        int q = carry2 + warp_excl_scan(v < n ? (n - 1) - popc(up) : 0, lane, &tot2);
        if (v < n)
          for (int j = 0; j < n; ++j)
            if (j != v && !((up >> j) & 1)) put_pair(o, pp + q++, nb + v, nb + j);
        carry2 += tot2;
      }
      pp += carry2;
    }
  }
  // ---- blocks 0_1 and 1_0
  for (int e = lane; e < n_e; e += 32) {
    const int a = s.ea[e], b = s.eb[e];
    put_pair(o, pp + 2 * e, nb + a, E0 + e);
    put_pair(o, pp + 2 * e + 1, nb + b, E0 + e);
    put_pair(o, pp + 2 * n_e + 2 * e, E0 + e, nb + a);
    put_pair(o, pp + 2 * n_e + 2 * e + 1, E0 + e, nb + b);
  }
  pp += 4ll * n_e;
  // ---- block 1_1: edges that share a triangle
  for (int e = lane; e < n_e; e += 32) {
    const int a = s.ea[e], b = s.eb[e];
    M m = s.cof[e];
    int64_t pos = pp + 2ll * s.cbase[e];
    while (m) {
      const int x = ffs0(m);
      m &= m - 1;
      // sorted triple (p, q, r); faces in gudhi boundary order (p,q), (p,r), (q,r) without e itself
      if (x < a) {          // (x, a, b): faces (x,a), (x,b), [a,b]
        put_pair(o, pos++, E0 + edge_index_of(s, x, a), E0 + e);
        put_pair(o, pos++, E0 + edge_index_of(s, x, b), E0 + e);
      } else if (x < b) {   // (a, x, b): faces (a,x), [a,b], (x,b)
        put_pair(o, pos++, E0 + edge_index_of(s, a, x), E0 + e);
        put_pair(o, pos++, E0 + edge_index_of(s, x, b), E0 + e);
      } else {              // (a, b, x): faces [a,b], (a,x), (b,x)
        put_pair(o, pos++, E0 + edge_index_of(s, a, x), E0 + e);
        put_pair(o, pos++, E0 + edge_index_of(s, b, x), E0 + e);
      }
    }
  }
  pp += 6ll * n_t;
  // ---- triangles: x_ind rows and blocks 1_2, 2_1
  for (int e = lane; e < n_e; e += 32) {
    const int p = s.ea[e], q = s.eb[e];
    M m = s.cof[e] & bits_above<M>(q);
    int64_t t = s.tbase[e];
    while (m) {
      const int r = ffs0(m);
      m &= m - 1;
      const int64_t f0 = E0 + e, f1 = E0 + edge_index_of(s, p, r), f2 = E0 + edge_index_of(s, q, r);
      put_pair(o, pp + 3 * t, f0, T0 + t);
      put_pair(o, pp + 3 * t + 1, f1, T0 + t);
      put_pair(o, pp + 3 * t + 2, f2, T0 + t);
      put_pair(o, pp + 3ll * n_t + 3 * t, T0 + t, f0);
      put_pair(o, pp + 3ll * n_t + 3 * t + 1, T0 + t, f1);
      put_pair(o, pp + 3ll * n_t + 3 * t + 2, T0 + t, f2);
      int v[3] = {p, q, r};
      float xr[3];
      frozenset_order(v, 3, xr);
      o.x_ind[(T0 + t) * 3] = xr[0]; o.x_ind[(T0 + t) * 3 + 1] = xr[1]; o.x_ind[(T0 + t) * 3 + 2] = xr[2];
      o.node_types[T0 + t] = 2;
      o.batch[T0 + t] = c;
      ++t;
    }
  }
}

inline int check_lift_desc(const csmpn_lift_desc* d) {
  if (!d || d->n_complexes < 0 || !d->vptr) return CSMPN_ERR_BAD_ARG;
  switch (d->mode) {
    case CSMPN_LIFT_RIPS:
      if (!d->points || d->point_dim <= 0) return CSMPN_ERR_BAD_ARG;
      break;
    case CSMPN_LIFT_CLIQUE:
    case CSMPN_LIFT_MOTION:
      if (!d->pptr || d->n_pairs < 0 || (d->n_pairs > 0 && !d->pairs)) return CSMPN_ERR_BAD_ARG;
      if (d->use_filters && (d->mode != CSMPN_LIFT_CLIQUE || !d->points || d->point_dim <= 0)) return CSMPN_ERR_BAD_ARG;
      break;
    case CSMPN_LIFT_KNN:
      if (!d->points || d->point_dim <= 0 || d->knn_k < 1) return CSMPN_ERR_BAD_ARG;
      break;
    case CSMPN_LIFT_FACETS:
      if (!d->fptr || !d->facets || d->facet_size < 2 || d->facet_size > kMaxV2) return CSMPN_ERR_BAD_ARG;
      break;
    default:
      return CSMPN_ERR_BAD_ARG;
  }
  if (d->max_dim < 1 || d->max_dim > 2) return CSMPN_ERR_UNSUPPORTED;
  if (d->max_vertices < 0 || d->max_vertices > kMaxV2) return CSMPN_ERR_UNSUPPORTED;
  return CSMPN_OK;
}

}  // namespace csmpn

using namespace csmpn;

extern "C" {

int csmpn_lift_count(const csmpn_lift_desc* desc, int32_t* counts, int64_t* node_ptr, int64_t* pair_ptr, int32_t* status,
                     csmpn_stream_t stream) {
  int st = check_lift_desc(desc);
  if (st) return st;
  if (!counts || !node_ptr || !pair_ptr || !status) return CSMPN_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  CSMPN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  if (desc->n_complexes > 0) {
    if (desc->max_vertices > kMaxV) {  // two-word masks, one complex per CTA (37 KB of tables per complex)
      lift_count_kernel<uint64_t, kMaxV2, 1><<<desc->n_complexes, 32, 0, s>>>(*desc, counts, status);
    } else {
      const int grid = (desc->n_complexes + kLiftWarps - 1) / kLiftWarps;
      lift_count_kernel<uint32_t, kMaxV, kLiftWarps><<<grid, 32 * kLiftWarps, 0, s>>>(*desc, counts, status);
    }
    CSMPN_LAUNCH_CHECK("lift_count");
  }
  lift_scan_kernel<<<1, 1024, 0, s>>>(*desc, counts, node_ptr, pair_ptr);
  CSMPN_LAUNCH_CHECK("lift_scan");
  return CSMPN_OK;
}

int csmpn_lift_fill(const csmpn_lift_desc* desc, const int64_t* node_ptr, const int64_t* pair_ptr, int64_t n_pairs_total,
                    int64_t* edge_index, float* x_ind, int64_t* node_types, int64_t* batch, csmpn_stream_t stream) {
  int st = check_lift_desc(desc);
  if (st) return st;
  if (!node_ptr || !pair_ptr || n_pairs_total < 0 || !x_ind || !node_types || !batch) return CSMPN_ERR_BAD_ARG;
  if (n_pairs_total > 0 && !edge_index) return CSMPN_ERR_BAD_ARG;
  if (desc->n_complexes == 0) return CSMPN_OK;
  LiftOut o{edge_index, edge_index + n_pairs_total, x_ind, node_types, batch};
  if (desc->max_vertices > kMaxV) {
    lift_fill_kernel<uint64_t, kMaxV2, 1><<<desc->n_complexes, 32, 0, (cudaStream_t)stream>>>(*desc, node_ptr, pair_ptr, o);
  } else {
    const int grid = (desc->n_complexes + kLiftWarps - 1) / kLiftWarps;
    lift_fill_kernel<uint32_t, kMaxV, kLiftWarps><<<grid, 32 * kLiftWarps, 0, (cudaStream_t)stream>>>(*desc, node_ptr, pair_ptr, o);
  }
  CSMPN_LAUNCH_CHECK("lift_fill");
  return CSMPN_OK;
}

}  // extern "C"
