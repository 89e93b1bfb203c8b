"""``SimplicialTransform`` / ``ManualTransform`` with the reference's interface (csmpn/data/modules/simplicial_data.py),
backed by the GPU lifter.  ``__call__(graph)`` handles one sample like the reference's ``pre_transform`` hook;
``lift(graphs)`` handles a list of samples in ONE launch and returns the collated batch object the models read
(edge_index, x_ind, x_ind_batch, x_ind_ptr, node_types, batch, ptr, zero-padded features).
"""
from __future__ import annotations

import torch

from .lifting import LIFT_CLIQUE, LIFT_FACETS, LIFT_KNN, LIFT_MOTION, LIFT_RIPS, lift_batch


class Data:
    """Minimal attribute container standing in for ``torch_geometric.data.Data`` (PyG is optional here)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return hasattr(self, k)

    @property
    def keys(self):
        return list(self.__dict__.keys())

    def to_dict(self):
        return dict(self.__dict__)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class SimplicialComplexData(Data):
    """Kept for API parity (simplicial_data.py:9-25); collation is done on the device by ``lift_batch``."""


_VERTEX_FEATURES = {"md17": ("loc", "vel", "charges"), "nba": ("pos", "vel"), "hulls": ("input",), "motion": ("loc", "vel")}


def _pad_vertex_features(feat_list, lb):
    """zero-padded per-simplex features: rows of vertices hold the vertex feature, edge / triangle rows are zero
    (simplicial_data.py:203-215, 224-247)."""
    x = torch.cat(feat_list, dim=0)          # concatenated where the samples live: ONE host-to-device copy per feature
    dev = lb.node_types.device
    x = x.to(dev, non_blocking=True)
    out = x.new_zeros((lb.num_nodes,) + tuple(x.shape[1:]))
    # vertex rows of complex c are the first n_c rows of its block: row = node_ptr[c] + local vertex id (no nonzero / sync)
    out.index_copy_(0, lb.vertex_rows, x)
    return out


class SimplicialTransform:
    def __init__(self, dim=2, dis: float = 2.0, label=None, edge_th=10000., tri_th=10000., molecule_type=None, knn_k=None):
        # knn_k (extension): build the kNN graph of the samples' first-frame positions on the GPU instead of reading
        # graph.edge_index (which the reference's dataset computes with torch_cluster on the CPU, data/md17.py:64)
        self.knn_k = knn_k
        self.dim = dim
        self.dis = dis
        self.label = label
        self.edge_th = edge_th
        self.tri_th = tri_th
        self.molecule_type = molecule_type

    # ---- batched path -------------------------------------------------------------------------------------
    def lift(self, graphs, device="cuda"):
        """Lift a list of samples in one launch; returns a collated batch (``Data``)."""
        if self.label not in ("md17", "nba", "hulls"):
            raise ValueError(f"Unknown dataset {self.label}.")
        nv = []
        if self.label == "hulls":
            from scipy.spatial import ConvexHull

            fac = [torch.as_tensor(ConvexHull(g.input.detach().cpu().numpy()).simplices).long()
                   if not hasattr(g, "facets") else g.facets for g in graphs]
            nv = [g.input.shape[0] for g in graphs]
            lb = lift_batch(LIFT_FACETS, nv, facets=torch.cat(fac, 0), facets_per_complex=[f.shape[0] for f in fac],
                            dim=self.dim, device=device)
        elif self.molecule_type == "aspirin":
            locs = [_loc(g) for g in graphs]
            nv = [l.shape[0] for l in locs]
            # the length / area filters of simplicial_lift (utils.py:181-200) are no-ops at the shipped 1e4
            filt = {}
            if self.edge_th < 1e4 or self.tri_th < 1e4:
                filt = dict(edge_th=self.edge_th, tri_th=self.tri_th)
            pts = torch.cat([_first_frame(l) for l in locs], 0) if (filt or self.knn_k) else None
            if self.knn_k:   # the kNN graph itself is built on the GPU (csmpn/data/md17.py:64): no edge_index needed
                lb = lift_batch(LIFT_KNN, nv, points=pts, knn_k=self.knn_k, device=device, **filt)
            else:
                lb = lift_batch(LIFT_CLIQUE, nv, pairs=torch.cat([g.edge_index for g in graphs], 1),
                                pairs_per_complex=[g.edge_index.shape[1] for g in graphs], points=pts, device=device, **filt)
        else:
            locs = [_loc(g) for g in graphs]
            nv = [l.shape[0] for l in locs]
            lb = lift_batch(LIFT_RIPS, nv, points=torch.cat([l.reshape(l.shape[0], -1) for l in locs], 0),
                            max_edge_length=self.dis, dim=self.dim, device=device)
        return _collate(lb, graphs, self.label)

    # ---- single-sample path (reference signature) -----------------------------------------------------------
    def __call__(self, graph):
        return self.lift([graph], device=_device_of(graph))


class ManualTransform:
    """The literal CMU-motion complex: 31 vertices, 12 edges, 4 triangles, 96 fixed pairs appended to the skeleton's
    0-0 pairs (simplicial_data.py:254-302)."""

    def __init__(self):
        self.name = "Manually adding triangles and edges to the graphs."
        self.num_edges = 12
        self.num_tris = 4
        self.num_nodes = 31
        self.dim = 2

    def lift(self, graphs, device="cuda"):
        lb = lift_batch(LIFT_MOTION, [31] * len(graphs), pairs=torch.cat([g.edge_index for g in graphs], 1),
                        pairs_per_complex=[g.edge_index.shape[1] for g in graphs], device=device)
        out = _collate(lb, graphs, "motion")
        out.pos = out.loc  # the reference stores the padded locations under ``pos`` (simplicial_data.py:324-327)
        return out

    def __call__(self, graph):
        return self.lift([graph], device=_device_of(graph))


def _loc(g):
    for name in ("init_pos", "loc", "pos"):
        if hasattr(g, name):
            return getattr(g, name)
    raise Exception("Graphs in datasets have to be specified with locations for constructing simplicial complexes.")


def _first_frame(loc):
    """vertex positions used for the filters / the kNN graph: [n, D], the first frame of [n, frames, D]"""
    return (loc[:, 0] if loc.dim() == 3 else loc).reshape(loc.shape[0], -1).float()


def _device_of(g):
    for v in vars(g).values():
        if torch.is_tensor(v) and v.is_cuda:
            return v.device
    return torch.device("cuda")


def _collate(lb, graphs, label):
    out = Data(edge_index=lb.edge_index, x_ind=lb.x_ind, node_types=lb.node_types, batch=lb.batch, ptr=lb.node_ptr,
               x_ind_batch=lb.batch, x_ind_ptr=lb.node_ptr, node_types_batch=lb.batch, node_types_ptr=lb.node_ptr,
               num_graphs=len(graphs))
    out.sizes = dict(lb.sizes)   # host ints: vertices / edges / triangles / pairs / simplices / complexes of this batch
    dev = lb.edge_index.device
    for name in _VERTEX_FEATURES[label]:
        if all(hasattr(g, name) for g in graphs):
            feats = [getattr(g, name) for g in graphs]
            if label == "md17" and name == "charges":
                frames = graphs[0].y.shape[1]
                feats = [f.unsqueeze(-1).repeat(1, frames).unsqueeze(-1) if f.dim() == 1 else f for f in feats]
            out[name] = _pad_vertex_features(feats, lb)
    if all(hasattr(g, "y") and g.y is not None for g in graphs):
        out.y = torch.cat([g.y for g in graphs], 0).to(dev, non_blocking=True)
    if all(hasattr(g, "target") for g in graphs):
        out.target = torch.stack([torch.as_tensor(g.target) for g in graphs]).to(dev)
    return out
