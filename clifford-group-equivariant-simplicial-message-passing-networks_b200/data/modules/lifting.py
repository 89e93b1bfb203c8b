"""Batched simplicial lifting on the GPU: Python side of csrc/lift.cu (C ABI: csmpn_lift_count / csmpn_lift_fill).

One call lifts a whole batch of complexes and returns them already collated (node offsets added, ``edge_index``
concatenated) -- what the reference obtains by running ``SimplicialTransform`` per sample on the CPU
(csmpn/data/modules/simplicial_data.py:40-103) and collating with PyG's DataLoader.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int32, c_int64, c_void_p

import torch

from ..._lib import CsmpnError, check, lib, ptr, require_cuda, stream_ptr

LIFT_RIPS, LIFT_CLIQUE, LIFT_FACETS, LIFT_MOTION, LIFT_KNN = 0, 1, 2, 3, 4
MAX_VERTICES = 64  # 33..64 vertices run the two-word-mask kernels (one complex per CTA)


class LiftDesc(ctypes.Structure):
    _fields_ = [
        ("mode", c_int32), ("n_complexes", c_int32), ("max_dim", c_int32), ("point_dim", c_int32),
        ("facet_size", c_int32), ("max_vertices", c_int32),
        ("max_edge_length", c_double), ("n_pairs", c_int64),
        ("vptr", c_void_p), ("points", c_void_p), ("pairs", c_void_p), ("pptr", c_void_p), ("facets", c_void_p),
        ("fptr", c_void_p),
        ("knn_k", c_int32), ("use_filters", c_int32), ("edge_th", ctypes.c_float), ("tri_th", ctypes.c_float),
    ]


class LiftedBatch:
    """Collated lift of ``n_complexes`` complexes.

    edge_index [2, E] int64 (global simplex ids; per complex the blocks 0_0,0_1,1_0,1_1,1_2,2_1 in that order),
    x_ind [N, 3] float32 (LOCAL vertex ids, zero padded), node_types [N] int64, batch [N] int64 (complex id per
    simplex = PyG's ``x_ind_batch`` / ``batch``), node_ptr / pair_ptr [n_complexes + 1] int64 (``ptr`` vectors),
    counts [n_complexes, 2] int32 (edges, triangles), n_vertices [n_complexes] int32.
    """

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return int(self.x_ind.shape[0])

    @property
    def vertex_rows(self):
        """[sum n_vertices] int64: row of every vertex in the collated simplex axis (the vertices of complex c are the
        first n_c rows of its block).  Device-side arithmetic only, no host sync."""
        rows = self.__dict__.get("_vertex_rows")
        if rows is None:
            nv = self.n_vertices.long()
            total = int(self.n_vertices_host.sum())
            first = torch.repeat_interleave(self.node_ptr[:-1], nv, output_size=total)
            v0 = torch.cumsum(nv, 0) - nv
            rows = first + torch.arange(total, device=nv.device) - torch.repeat_interleave(v0, nv, output_size=total)
            self._vertex_rows = rows
        return rows

    def block_sizes(self, c: int, single: bool):
        """pair counts of the six adjacency blocks of complex c (host ints)."""
        n = int(self.n_vertices_host[c])
        ne, nt = (int(v) for v in self.counts_host[c])
        b00 = 2 * ne + ((n * (n - 1) - ne) if single else 0)
        return {"0_0": b00, "0_1": 2 * ne, "1_0": 2 * ne, "1_1": 6 * nt, "1_2": 3 * nt, "2_1": 3 * nt}


def _i32(t, device):
    return t.to(device=device, dtype=torch.int32).contiguous()


def lift_batch(mode: int, n_vertices, *, points=None, max_edge_length=0.0, pairs=None, pairs_per_complex=None, facets=None,
               facets_per_complex=None, dim: int = 2, device=None, knn_k=None, edge_th=None, tri_th=None) -> LiftedBatch:
    """Lift a batch.  ``n_vertices``: int tensor / list [n_complexes].  mode RIPS: ``points`` [sum n, D] fp32.
    mode CLIQUE / MOTION: ``pairs`` [2, P] int64 local ids concatenated over complexes + ``pairs_per_complex``.
    mode KNN: CLIQUE whose graph is ``knn_graph(points, knn_k)`` (csmpn/data/md17.py:64), built inside the kernel.
    CLIQUE / KNN with ``edge_th`` / ``tri_th`` (and ``points``): the length / area filters of utils.py:181-200.
    mode FACETS: ``facets`` [F, k] int64 local ids + ``facets_per_complex``."""
    anchor = points if points is not None else (pairs if pairs is not None else facets)
    if device is None:
        device = anchor.device
    device = torch.device(device)
    if device.type != "cuda":
        raise CsmpnError(f"csmpn_b200.lift_batch: expected a CUDA device (got {device}); this package has no CPU path")
    nv = torch.as_tensor(n_vertices, dtype=torch.int64)
    ncx = int(nv.numel())
    if ncx and int(nv.max()) > MAX_VERTICES:
        raise ValueError(f"lift_batch supports at most {MAX_VERTICES} vertices per complex (got {int(nv.max())})")

    def make_ptr(counts):
        c = torch.as_tensor(counts, dtype=torch.int64)
        return _i32(torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(c, 0)]), device)

    d = LiftDesc()
    d.mode, d.n_complexes, d.max_dim = mode, ncx, dim
    d.max_vertices = int(nv.max()) if ncx else 0  # > 32 selects the two-word-mask kernels
    vptr = make_ptr(nv)
    d.vptr = vptr.data_ptr()
    keep = [vptr]
    if mode == LIFT_RIPS:
        pts = points.to(device=device, dtype=torch.float32).contiguous()
        if pts.dim() != 2 or pts.shape[0] != int(nv.sum()):
            raise ValueError(f"points must be [sum(n_vertices), D]; got {tuple(pts.shape)} for {int(nv.sum())} vertices")
        d.points, d.point_dim, d.max_edge_length = pts.data_ptr(), pts.shape[1], float(max_edge_length)
        keep.append(pts)
    elif mode == LIFT_KNN:
        if points is None or knn_k is None or int(knn_k) < 1:
            raise ValueError("LIFT_KNN needs points [sum n, D] and knn_k >= 1")
        pts = points.to(device=device, dtype=torch.float32).contiguous()
        if pts.dim() != 2 or pts.shape[0] != int(nv.sum()):
            raise ValueError(f"points must be [sum(n_vertices), D]; got {tuple(pts.shape)} for {int(nv.sum())} vertices")
        d.points, d.point_dim, d.knn_k = pts.data_ptr(), pts.shape[1], int(knn_k)
        keep.append(pts)
    elif mode in (LIFT_CLIQUE, LIFT_MOTION):
        pr = pairs.to(device=device, dtype=torch.int64).contiguous()
        if pr.dim() != 2 or pr.shape[0] != 2:
            raise ValueError(f"pairs must be [2, P]; got {tuple(pr.shape)}")
        pptr = make_ptr(pairs_per_complex)
        d.pairs, d.n_pairs, d.pptr = pr.data_ptr(), pr.shape[1], pptr.data_ptr()
        keep += [pr, pptr]
        if mode == LIFT_MOTION and ncx and not bool((nv == 31).all()):
            raise ValueError("the motion template (ManualTransform) has exactly 31 vertices per complex")
    elif mode == LIFT_FACETS:
        fc = facets.to(device=device, dtype=torch.int64).contiguous()
        fptr = make_ptr(facets_per_complex)
        d.facets, d.facet_size, d.fptr = fc.data_ptr(), fc.shape[1], fptr.data_ptr()
        keep += [fc, fptr]
    else:
        raise ValueError(f"unknown lifting mode {mode}")

    if edge_th is not None or tri_th is not None:
        if mode not in (LIFT_CLIQUE, LIFT_KNN) or points is None:
            raise ValueError("edge_th / tri_th filters apply to the clique lifts and need the vertex positions (points)")
        if mode == LIFT_CLIQUE:
            pts = points.to(device=device, dtype=torch.float32).contiguous()
            if pts.dim() != 2 or pts.shape[0] != int(nv.sum()):
                raise ValueError(f"points must be [sum(n_vertices), D]; got {tuple(pts.shape)}")
            d.points, d.point_dim = pts.data_ptr(), pts.shape[1]
            keep.append(pts)
        d.use_filters = 1
        d.edge_th = float("inf") if edge_th is None else float(edge_th)
        d.tri_th = float("inf") if tri_th is None else float(tri_th)
    counts = torch.empty((ncx, 2), dtype=torch.int32, device=device)
    node_ptr = torch.empty(ncx + 1, dtype=torch.int64, device=device)
    pair_ptr = torch.empty(ncx + 1, dtype=torch.int64, device=device)
    status = torch.empty(1, dtype=torch.int32, device=device)
    s = stream_ptr(device)
    check(lib().csmpn_lift_count(ctypes.byref(d), ptr(counts), ptr(node_ptr), ptr(pair_ptr), ptr(status), s), "lift_count")
    # the one host sync: output sizes (+ the per-dimension totals, which shape-padded batches need: data/padding.py)
    csum = counts.long().sum(0) if ncx else torch.zeros(2, dtype=torch.int64, device=device)
    totals = torch.stack([node_ptr[-1], pair_ptr[-1], status[0].long(), csum[0], csum[1]]).cpu()
    n_total, e_total, bad, n_edges_total, n_tris_total = (int(v) for v in totals)
    if bad:
        raise ValueError("lift_batch: a complex has an unsupported vertex count")
    edge_index = torch.empty((2, e_total), dtype=torch.int64, device=device)
    x_ind = torch.empty((n_total, 3), dtype=torch.float32, device=device)
    node_types = torch.empty(n_total, dtype=torch.int64, device=device)
    batch = torch.empty(n_total, dtype=torch.int64, device=device)
    check(lib().csmpn_lift_fill(ctypes.byref(d), ptr(node_ptr), ptr(pair_ptr), e_total, ptr(edge_index), ptr(x_ind),
                                ptr(node_types), ptr(batch), s), "lift_fill")
    out = LiftedBatch(edge_index=edge_index, x_ind=x_ind, node_types=node_types, batch=batch, node_ptr=node_ptr,
                      pair_ptr=pair_ptr, counts=counts, n_vertices=vptr[1:] - vptr[:-1], mode=mode, _keep=keep)
    out.n_vertices_host = nv
    out.sizes = {"vertices": int(nv.sum()) if ncx else 0, "edges": n_edges_total, "triangles": n_tris_total, "pairs": e_total,
                 "simplices": n_total, "complexes": ncx}
    out._counts_host = None
    return out


def _counts_host(self):
    if self._counts_host is None:
        self._counts_host = self.counts.cpu()
    return self._counts_host


LiftedBatch.counts_host = property(_counts_host)
