"""Pure-Python restatement of the simplicial lifting (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates csmpn/data/modules/utils.py (lifting, indices, adjacencies) and simplicial_data.py (merging into one
``edge_index`` / ``x_ind`` / ``node_types``) of the reference.  The reference builds its complexes with
gudhi 3.9.0 (``environment.yml:18``), which is neither vendored under /root/reference nor installable here, so
``SimplexTree`` / ``RipsComplex`` below restate the published behaviour of the pieces the reference calls
(utils.py:30,44,67,71-72,75,128-130,179-200,230-241):

  insert(s)              adds s and all its faces (vertices kept sorted)
  get_simplices()        depth-first traversal of the prefix tree, children before their parent
                         ("Complex_simplex_iterator"); inside one dimension this is lexicographic order
  get_boundaries(s)      facets of s, dropping the largest vertex first ... the smallest last
                         ("Boundary_simplex_iterator")
  get_cofaces(s, 1)      codimension-1 cofaces in lexicographic order (rec_coface walks the tree in label order)
  RipsComplex(points, max_edge_length).create_simplex_tree(max_dimension)
                         clique complex of {|p_i - p_j| <= max_edge_length} (float64 distances) up to max_dimension

PARITY UNPINNED against real gudhi for those traversal orders; everything downstream of them is pinned against
the reference's own utils.py / simplicial_data.py executed over this stand-in (tests/golden/make_golden.py).
"""
from __future__ import annotations

import itertools
import math

import torch


# =================================================================================== gudhi stand-ins
class SimplexTree:
    def __init__(self):
        self._simplices = {}  # sorted vertex tuple -> filtration

    def insert(self, simplex, filtration=0.0):
        s = tuple(sorted(int(v) for v in simplex))
        new = False
        for r in range(1, len(s) + 1):
            for face in itertools.combinations(s, r):
                if face not in self._simplices:
                    self._simplices[face] = float(filtration)
                    new = True
        return new

    def _children(self, prefix):
        n = len(prefix)
        return sorted(s for s in self._simplices if len(s) == n + 1 and s[:n] == prefix)

    def get_simplices(self):
        out = []

        def visit(node):
            for ch in self._children(node):
                visit(ch)
            if node:
                out.append((list(node), self._simplices[node]))

        visit(())
        return iter(out)

    def get_boundaries(self, simplex):
        s = tuple(sorted(int(v) for v in simplex))
        out = []
        if len(s) > 1:
            for drop in range(len(s) - 1, -1, -1):
                face = s[:drop] + s[drop + 1:]
                out.append((list(face), self._simplices[face]))
        return iter(out)

    def get_cofaces(self, simplex, codimension):
        s = set(int(v) for v in simplex)
        n = len(s) + codimension
        return [(list(t), f) for t, f in sorted(self._simplices.items()) if len(t) == n and s.issubset(t)]

    def num_simplices(self):
        return len(self._simplices)


class RipsComplex:
    def __init__(self, points=None, max_edge_length=float("inf")):
        self.points = [list(map(float, p)) for p in points]
        self.max_edge_length = float(max_edge_length)

    def create_simplex_tree(self, max_dimension=1):
        st = SimplexTree()
        n = len(self.points)
        for v in range(n):
            st.insert([v])
        adj = [set() for _ in range(n)]
        for i in range(n):
            for j in range(i + 1, n):
                d = math.sqrt(sum((a - b) ** 2 for a, b in zip(self.points[i], self.points[j])))
                if d <= self.max_edge_length:
                    adj[i].add(j)
                    adj[j].add(i)
                    st.insert([i, j], d)
        if max_dimension >= 2:
            for i in range(n):
                for j in sorted(adj[i]):
                    if j <= i:
                        continue
                    for k in sorted(adj[i] & adj[j]):
                        if k > j:
                            st.insert([i, j, k])
        if max_dimension >= 3:
            raise NotImplementedError("stand-in expands to dimension 2 (all the reference uses)")
        return st


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    """torch_cluster 1.6.1 ``knn_graph`` restated: row 0 = neighbour j, row 1 = centre i, k nearest by Euclidean
    distance, nearest first, self excluded (csmpn/data/md17.py:64, nba.py:48)."""
    assert batch is None and flow == "source_to_target"
    n = x.shape[0]
    d = torch.cdist(x.double(), x.double())
    if not loop:
        d.fill_diagonal_(float("inf"))
    kk = min(int(k), n if loop else n - 1)
    order = torch.argsort(d, dim=1, stable=True)[:, :kk]
    centre = torch.arange(n).unsqueeze(1).expand(n, kk)
    return torch.stack([order.reshape(-1), centre.reshape(-1)]).long()


# =================================================================================== lifting restated
def _dict_indices(simplices_in_order):
    """utils.py:37-51 / 285-319: running index per dimension in traversal order."""
    indices = {}
    for s in simplices_in_order:
        d = len(s) - 1
        tab = indices.setdefault(d, {})
        fs = frozenset(s)
        if fs not in tab:
            tab[fs] = len(tab)
    return indices


def _features(indices):
    """utils.py:53-61 / 378-388: x_dict[d][idx] = list(frozenset) -- vertex order is CPython set order."""
    x = {}
    for d in sorted(indices):
        rows = [None] * len(indices[d])
        for fs, idx in indices[d].items():
            rows[idx] = list(fs)
        x[d] = torch.tensor(rows, dtype=torch.long).reshape(len(rows), d + 1)
    return x


def _adjacencies(st, indices, simplices_in_order, extra_zero_zero):
    """utils.py:63-103 (``_single``, extra_zero_zero=True) and utils.py:322-375 (extra_zero_zero=False)."""
    adj = {}
    for s in simplices_in_order:
        d = len(s) - 1
        fs = frozenset(s)
        if fs not in indices.get(d, {}):
            continue
        si = indices[d][fs]
        for coface, _ in st.get_cofaces(s, 1):
            for cb, _ in st.get_boundaries(coface):
                if frozenset(cb) != fs:
                    adj.setdefault(f"{d}_{d}", []).append([indices[d][frozenset(cb)], si])
        for b, _ in st.get_boundaries(s):
            if frozenset(b) in indices.get(d - 1, {}):
                adj.setdefault(f"{d - 1}_{d}", []).append([indices[d - 1][frozenset(b)], si])
    if extra_zero_zero:
        n0 = len(indices[0])
        present = [s for s in simplices_in_order if len(s) == 2]
        for i in range(n0):
            for j in range(n0):
                if i != j and [i, j] not in present:
                    adj.setdefault("0_0", []).append([i, j])
    return {k: torch.tensor(v, dtype=torch.long).T.contiguous() for k, v in adj.items()}


def lift_from_tree(st, single: bool):
    """Common tail of rips_lift / simplicial_lift_hulls (single=True) and simplicial_lift (single=False)."""
    order = [s for s, _ in st.get_simplices()]
    if not single:
        # utils.py:250-274: triangles are kept only if one of their edges is recorded (always true here because
        # insert() adds all faces); vertices, then edges, then triangles, each in traversal order
        order_idx = [s for s in order if len(s) == 1] + [s for s in order if len(s) == 2] + [s for s in order if len(s) == 3]
        indices = _dict_indices(order_idx)
    else:
        indices = _dict_indices(order)
    adj = _adjacencies(st, indices, order, extra_zero_zero=single)
    return _features(indices), adj


def rips_lift_ref(points, dim, dis):
    """utils.py:106-136"""
    st = RipsComplex(points=[list(map(float, p)) for p in points], max_edge_length=dis).create_simplex_tree(max_dimension=dim)
    return lift_from_tree(st, single=True)


def clique_lift_ref(n_vertices, edge_index, loc=None, edge_th=None, tri_th=None):
    """utils.py:151-207: clique complex (<= triangles) of the kNN graph.  With ``loc`` [n, 3] the filters of
    utils.py:181-200 apply: a graph edge is inserted iff its length <= edge_th, a 3-clique iff its area
    (utils.py:139-148, per-triangle formula) <= tri_th -- and SimplexTree.insert adds every face of an inserted
    triangle, so an edge longer than edge_th still enters through a kept triangle.  Without ``loc``: the shipped
    thresholds of 1e4, i.e. no filtering."""
    nbr = [set() for _ in range(n_vertices)]
    for a, b in edge_index.t().tolist():
        if a != b:
            nbr[a].add(b)
            nbr[b].add(a)

    def length(a, b):
        return float(torch.norm(loc[a] - loc[b]))

    def area(a, b, c):
        return float(0.5 * torch.linalg.norm(torch.linalg.cross(loc[b] - loc[a], loc[c] - loc[a])))

    st = SimplexTree()
    for v in range(n_vertices):
        st.insert([v])
    for a in range(n_vertices):
        for b in sorted(nbr[a]):
            if b > a and (loc is None or edge_th is None or length(a, b) <= edge_th):
                st.insert([a, b])
    for a in range(n_vertices):
        for b in sorted(nbr[a]):
            if b <= a:
                continue
            for c in sorted(nbr[a] & nbr[b]):
                if c > b and (loc is None or tri_th is None or area(a, b, c) <= tri_th):
                    st.insert([a, b, c])
    return lift_from_tree(st, single=False)


def hull_faces_lift_ref(n_vertices, facets, dim):
    """utils.py:210-248: every <= dim-face of every Qhull facet."""
    st = SimplexTree()
    for v in range(n_vertices):
        st.insert([v])
    for k in range(1, dim + 1):
        faces = set()
        for f in facets:
            for sub in itertools.combinations([int(v) for v in f], k + 1):
                faces.add(tuple(sorted(sub)))
        for face in faces:
            st.insert(face)
    return lift_from_tree(st, single=True)


def merge_ref(x_dict, adj, max_dim=2):
    """simplicial_data.py:105-175,218-222: reversed (coboundary) blocks, per-dimension offsets, block order
    0_0,0_1,1_0,1_1,1_2,2_1; x_ind float32 [N,3] zero padded; node_types."""
    adj = dict(adj)
    for d in range(max_dim):
        if f"{d}_{d + 1}" in adj:
            adj[f"{d + 1}_{d}"] = adj[f"{d}_{d + 1}"][[1, 0]].clone()
    offs = [0]
    for d in x_dict:
        offs.append(offs[-1] + len(x_dict[d]))
    blocks = []
    for ds in x_dict:
        for dt in x_dict:
            key = f"{ds}_{dt}"
            if key in adj:
                e = torch.zeros_like(adj[key])
                e[0] = adj[key][0] + offs[ds]
                e[1] = adj[key][1] + offs[dt]
                blocks.append(e)
    edge_index = torch.cat(blocks, dim=-1)
    n = offs[-1]
    x_ind = torch.zeros((n, 3))
    node_types = torch.zeros(n, dtype=torch.long)
    for d in x_dict:
        x_ind[offs[d]: offs[d + 1], : d + 1] = x_dict[d]
        node_types[offs[d]: offs[d + 1]] = d
    return edge_index, x_ind, node_types


# motion: ManualTransform's literal complex (simplicial_data.py:263-296)
MOTION_EDGES = [[6, 7], [7, 8], [6, 8], [1, 2], [2, 3], [1, 3], [24, 25], [25, 26], [24, 26], [22, 23], [21, 22], [21, 23]]
MOTION_TRIS = [[6, 7, 8], [1, 2, 3], [24, 25, 26], [21, 22, 23]]


def motion_manual_ref(base_edge_index):
    """edge_index = cat(skeleton 0-0 pairs, 2->1 / 1->2, 1->0 / 0->1, 1<->1) for the fixed CMU complex."""
    e12_tri = [43 + t for t in range(4) for _ in range(3)]
    e12_edge = list(range(31, 43))
    dim1_dim2 = torch.tensor([e12_tri + e12_edge, e12_edge + e12_tri])
    edge_ids = [31 + e for e in range(12) for _ in range(2)]
    verts = [v for e in MOTION_EDGES for v in e]
    dim1_dim0 = torch.tensor([edge_ids + verts, verts + edge_ids])
    a, b = [], []
    for t in range(4):
        es = [31 + 3 * t + q for q in range(3)]
        for x in es:
            for y in es:
                if x != y:
                    a.append(x), b.append(y)
    dim1_dim1 = torch.tensor([a, b])
    edge_index = torch.cat((base_edge_index, dim1_dim2, dim1_dim0, dim1_dim1), dim=-1)
    x_ind = torch.zeros((47, 3))
    x_ind[:31, :1] = torch.arange(31).unsqueeze(1).float()
    x_ind[31:43, :2] = torch.tensor(MOTION_EDGES).float()
    x_ind[43:47, :3] = torch.tensor(MOTION_TRIS).float()
    node_types = torch.tensor([0] * 31 + [1] * 12 + [2] * 4)
    return edge_index, x_ind, node_types
