"""Lifting entry points with the reference's names (csmpn/data/modules/utils.py), running on the GPU lifter.

``rips_lift`` / ``simplicial_lift`` / ``simplicial_lift_hulls`` take one graph object (anything with the attributes
the reference reads: ``init_pos`` | ``loc`` | ``pos``, ``edge_index``, ``input``) and return ``(x_dict, adj)`` exactly
like the reference (utils.py:106-136, 151-207, 210-248): ``x_dict[d]`` int64 ``[n_d, d+1]`` with the vertices of a
simplex in CPython-frozenset order, ``adj["s_t"]`` int64 ``[2, n]`` with per-dimension indices.  For whole batches
use ``lifting.lift_batch`` directly -- one launch for all complexes.

Qhull (``scipy.spatial.ConvexHull``, utils.py:219-221) stays on the host: it is a third-party geometry routine, not
part of the lifting; its facets are the input of the GPU path.
"""
from __future__ import annotations

import torch

from . import lifting
from .lifting import LIFT_CLIQUE, LIFT_FACETS, LIFT_RIPS, lift_batch


def _locations(graph):
    for name in ("init_pos", "loc", "pos"):
        if hasattr(graph, name):
            return getattr(graph, name)
    raise Exception("Graphs in datasets have to be specified with locations for constructing simplicial complexes.")


def split_single(lb: "lifting.LiftedBatch", single: bool, c: int = 0):
    """(x_dict, adj) of complex ``c`` of a lifted batch, in the reference's per-sample format."""
    n0, p0 = int(lb.node_ptr[c]), int(lb.pair_ptr[c])
    n = int(lb.n_vertices_host[c])
    ne, nt = (int(v) for v in lb.counts_host[c])
    sizes = [n, ne, nt]
    offs = [0, n, n + ne, n + ne + nt]
    x_dict, adj = {}, {}
    for d in range(3):
        if sizes[d] or d == 0:
            x_dict[d] = lb.x_ind[n0 + offs[d]: n0 + offs[d + 1], : d + 1].long()
    pos = p0
    for key, cnt in lb.block_sizes(c, single).items():
        if cnt:
            s, t = int(key[0]), int(key[2])
            blk = lb.edge_index[:, pos: pos + cnt].clone()
            blk[0] -= n0 + offs[s]
            blk[1] -= n0 + offs[t]
            adj[key] = blk
        pos += cnt
    return x_dict, adj


def rips_lift(graph, dim: int, dis: float):
    loc = _locations(graph)
    lb = lift_batch(LIFT_RIPS, [loc.shape[0]], points=loc.reshape(loc.shape[0], -1), max_edge_length=dis, dim=dim)
    x_dict, adj = split_single(lb, single=True)
    return x_dict, {k: v for k, v in adj.items() if k in ("0_0", "0_1", "1_1", "1_2")}


def triangle_area(vertex1, vertex2, vertex3):
    """0.5 |(v2 - v1) x (v3 - v1)| per row (utils.py:139-148).  The reference calls ``torch.cross`` without ``dim``, which
    picks the FIRST axis of size 3 -- the wrong one when exactly three triangles are passed; this is the per-row formula
    (and what the GPU lifter evaluates)."""
    return 0.5 * torch.linalg.norm(torch.linalg.cross(vertex2 - vertex1, vertex3 - vertex1, dim=-1), dim=-1)


def simplicial_lift(graph, edge_th=10000, tri_th=10000):
    loc = _locations(graph)
    ei = graph.edge_index
    filt = {}
    if edge_th < 1e4 or tri_th < 1e4:  # utils.py:181-200; no-ops at the shipped thresholds
        pts = (loc[:, 0] if loc.dim() == 3 else loc).reshape(loc.shape[0], -1).float()
        filt = dict(points=pts, edge_th=edge_th, tri_th=tri_th)
    lb = lift_batch(LIFT_CLIQUE, [loc.shape[0]], pairs=ei, pairs_per_complex=[ei.shape[1]], device=ei.device, **filt)
    x_dict, adj = split_single(lb, single=False)
    return x_dict, {k: v for k, v in adj.items() if k in ("0_0", "0_1", "1_1", "1_2")}


def simplicial_lift_hulls(graph, dim: int, facets=None):
    pts = graph.input
    if facets is None:
        from scipy.spatial import ConvexHull

        facets = torch.as_tensor(ConvexHull(pts.detach().cpu().numpy()).simplices).long()
    lb = lift_batch(LIFT_FACETS, [pts.shape[0]], facets=facets, facets_per_complex=[facets.shape[0]], dim=dim,
                    device=pts.device)
    x_dict, adj = split_single(lb, single=True)
    return x_dict, {k: v for k, v in adj.items() if k in ("0_0", "0_1", "1_1", "1_2")}
