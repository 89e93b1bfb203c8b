"""Shape-padded batches: a stream of DIFFERENT complexes replayed from ONE CUDA graph.

MD17 / NBA batches change their simplex and pair counts from step to step (the kNN / Rips complexes differ from sample to
sample), while a CUDA graph is captured for fixed shapes.  A ``Bucket`` fixes the per-dimension simplex counts and the
pair count; ``pad_to_bucket`` appends ONE dummy complex to a collated batch (``SimplicialTransform.lift``) that absorbs
the difference:

* ``n_dummy_vertices`` vertices (as many as a real complex has, so per-graph reshapes of the models keep working) with
  all-zero features, then the missing edges (local vertices 0, 1) and triangles (0, 1, 2);
* the missing pairs are self-pairs of the dummy simplices, dealt round-robin.

Dummy simplices exchange messages only with dummy simplices, are pooled into their own graph (index = number of real
graphs) and are excluded from the loss (``n_real_graphs``): outputs and every parameter gradient of the real complexes
are unchanged (tests/test_models.py::test_padded_stream_step_*).  Everything here is device-side ``cat`` / ``copy_`` on
host-known sizes (``batch.sizes``, filled by the lifter's one size sync): no additional synchronisation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .modules.simplicial_data import Data

NODE_ALIGNED = ("loc", "vel", "charges", "pos", "input")   # zero-padded per-simplex features (simplicial_data.py:203-247)
VERTEX_ALIGNED = ("y",)                                      # one row per vertex
GRAPH_ALIGNED = ("target",)                                  # one row per complex
_BATCH_KEYS = ("batch", "x_ind_batch", "node_types_batch")
_PTR_KEYS = ("ptr", "x_ind_ptr", "node_types_ptr")


class BucketOverflow(ValueError):
    """the batch has more edges / triangles / pairs than the bucket holds (capture a larger bucket)"""


@dataclass(frozen=True)
class Bucket:
    vertices: int          # real vertices of a batch (fixed: complexes x vertices per complex)
    n_dummy_vertices: int  # vertices of the dummy complex
    edges: int             # padded totals, dummy complex included
    triangles: int
    pairs: int
    complexes: int         # real complexes per batch

    @property
    def simplices(self) -> int:
        return self.vertices + self.n_dummy_vertices + self.edges + self.triangles

    def counts(self):
        """simplices per dimension of a padded batch (host ints): what the models' index caches are sized with"""
        return (self.vertices + self.n_dummy_vertices, self.edges, self.triangles)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def make_bucket(sizes_list, margin: float = 1.02, multiple: int = 64) -> Bucket:
    """Bucket covering every ``batch.sizes`` in ``sizes_list`` (e.g. a pass over the sampler's batches, or the first few
    batches) with ``margin`` head room; all batches must have the same number of complexes and vertices."""
    sizes_list = list(sizes_list)
    v, ncx = sizes_list[0]["vertices"], sizes_list[0]["complexes"]
    if any(s["vertices"] != v or s["complexes"] != ncx for s in sizes_list):
        raise ValueError("make_bucket: batches of one bucket need the same complex and vertex counts")
    nvd = v // ncx if (ncx and v % ncx == 0) else 3
    nvd = max(nvd, 3)
    grow = lambda key, extra: _round_up(int(math.ceil(max(s[key] for s in sizes_list) * margin)) + extra, multiple)
    return Bucket(vertices=v, n_dummy_vertices=nvd, edges=grow("edges", 1), triangles=grow("triangles", 1), pairs=grow("pairs", 1),
                  complexes=ncx)


def pad_to_bucket(batch: Data, bucket: Bucket, out: Data | None = None) -> Data:
    """Collated batch -> the same batch plus one dummy complex, with exactly ``bucket`` sizes.  With ``out`` (a padded batch
    made earlier) the tensors of ``out`` are overwritten in place -- what a captured CUDA graph needs -- else new ones are made."""
    sz = batch.sizes
    if sz["vertices"] != bucket.vertices or sz["complexes"] != bucket.complexes:
        raise ValueError(f"pad_to_bucket: batch has {sz['complexes']} complexes / {sz['vertices']} vertices, bucket "
                         f"{bucket.complexes} / {bucket.vertices}")
    pe, pt, pp = bucket.edges - sz["edges"], bucket.triangles - sz["triangles"], bucket.pairs - sz["pairs"]
    if min(pe, pt, pp) < 0:
        raise BucketOverflow(f"batch sizes {sz} exceed the bucket {bucket}")
    nvd = bucket.n_dummy_vertices
    n, n_pad = sz["simplices"], bucket.simplices
    tail = nvd + pe + pt
    assert n + tail == n_pad, (n, tail, n_pad)
    dev = batch.edge_index.device
    B = bucket.complexes
    res = out if out is not None else Data()

    def put(key, real, tail_tensor):
        """res[key] = cat(real, tail_tensor), written in place when res already holds the tensor"""
        cur = getattr(res, key, None) if out is not None else None
        if cur is None:
            setattr(res, key, torch.cat([real, tail_tensor.to(real.dtype)], 0) if tail_tensor.shape[0] else real.clone())
        else:
            k = real.shape[0]
            cur[:k].copy_(real)
            if tail_tensor.shape[0]:
                cur[k:].copy_(tail_tensor)

    # ---- structure
    types_tail = torch.cat([torch.zeros(nvd, dtype=torch.int64, device=dev), torch.ones(pe, dtype=torch.int64, device=dev),
                            torch.full((pt,), 2, dtype=torch.int64, device=dev)])
    put("node_types", batch.node_types, types_tail)
    xi = torch.zeros((tail, 3), dtype=batch.x_ind.dtype, device=dev)
    xi[:nvd, 0] = torch.arange(nvd, device=dev, dtype=batch.x_ind.dtype)   # vertex i: (i, 0, 0)
    xi[nvd:, 1] = 1                                                         # edges (0, 1, 0), triangles (0, 1, 2)
    xi[nvd + pe:, 2] = 2
    put("x_ind", batch.x_ind, xi)
    put("batch", batch.batch, torch.full((tail,), B, dtype=torch.int64, device=dev))
    put("ptr", batch.ptr, (batch.ptr[-1:] + tail))
    for k in _BATCH_KEYS[1:]:
        setattr(res, k, res.batch)
    for k in _PTR_KEYS[1:]:
        setattr(res, k, res.ptr)
    # self-pairs of the dummy simplices, dealt round-robin over all of them (rows n .. n_pad-1): one dummy receiver with
    # thousands of pairs would serialise the counting sort and the per-receiver reductions (measured: step 2x slower)
    pad_pairs = (n + torch.arange(pp, device=dev, dtype=torch.int64) % tail).unsqueeze(0).expand(2, -1)
    cur = getattr(res, "edge_index", None) if out is not None else None
    if cur is None:
        res.edge_index = torch.cat([batch.edge_index, pad_pairs], 1)
    else:
        cur[:, : sz["pairs"]].copy_(batch.edge_index)
        if pp:
            cur[:, sz["pairs"]:].copy_(pad_pairs)
    # ---- features
    for key in NODE_ALIGNED:
        if hasattr(batch, key) and torch.is_tensor(getattr(batch, key)):
            t = getattr(batch, key)
            put(key, t, t.new_zeros((tail,) + tuple(t.shape[1:])))
    for key in VERTEX_ALIGNED:
        if hasattr(batch, key) and torch.is_tensor(getattr(batch, key)):
            t = getattr(batch, key)
            # rows per complex: one per vertex (md17: 21 atoms), or fewer (nba: the 10 players of 11 vertices have targets)
            per = t.shape[0] // B if (B and t.shape[0] % B == 0) else nvd
            put(key, t, t.new_zeros((per,) + tuple(t.shape[1:])))
    for key in GRAPH_ALIGNED:
        if hasattr(batch, key) and torch.is_tensor(getattr(batch, key)):
            t = getattr(batch, key)
            put(key, t, t.new_zeros((1,) + tuple(t.shape[1:])))
    res.num_graphs = B + 1
    res.n_real_graphs = B
    res.pad_counts = bucket.counts()
    res._csmpn_dynamic = True   # per-batch index caches of the models are recomputed every forward (inside the graph)
    res.sizes = {"vertices": bucket.vertices + nvd, "edges": bucket.edges, "triangles": bucket.triangles, "pairs": bucket.pairs,
                 "simplices": n_pad, "complexes": B + 1}
    res.real_sizes = dict(sz)
    return res
