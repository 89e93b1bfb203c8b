"""Stand-ins that let the unmodified reference import in a container without PyG / torch_scatter / gudhi.

TEST INFRASTRUCTURE.  Restates (does not copy) the third-party behaviour the reference leans on:
  * torch_geometric 2.3.0 ``MessagePassing.propagate`` for flow source_to_target (gather by edge_index[0] -> *_j,
    edge_index[1] -> *_i; aggregate at edge_index[1] with sum / mean(count clamped to 1); then ``update``),
  * ``global_mean_pool``; ``Data`` as an attribute bag; gudhi ``SimplexTree`` / ``RipsComplex`` (oracle.lift_ref).
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch

def _default_root() -> str:
    """/root/reference in the build container; on the GPU box the verbatim copy oracle/make_ref.py left in oracle/_ref"""
    if os.path.isdir("/root/reference/csmpn"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = os.environ.get("CSMPN_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "csmpn"))


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        msg_params = list(inspect.signature(self.message).parameters)
        upd_params = list(inspect.signature(self.update).parameters)[1:]
        margs = {}
        for name in msg_params:
            if name.endswith("_j"):
                margs[name] = kwargs[name[:-2]].index_select(0, edge_index[0])
            elif name.endswith("_i"):
                margs[name] = kwargs[name[:-2]].index_select(0, edge_index[1])
            else:
                margs[name] = kwargs.get(name)
        out = self.message(**margs)
        n = kwargs["h"].size(0)
        agg = torch.zeros((n,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
        agg.index_add_(0, edge_index[1], out)
        if self.aggr == "mean":
            cnt = torch.zeros(n, dtype=out.dtype, device=out.device)
            cnt.index_add_(0, edge_index[1], torch.ones(edge_index.shape[1], dtype=out.dtype, device=out.device))
            agg = agg / cnt.clamp(min=1).view(-1, *([1] * (out.dim() - 1)))
        elif self.aggr not in ("sum", "add"):
            raise NotImplementedError(self.aggr)
        return self.update(agg, **{p: kwargs.get(p) for p in upd_params})


def global_mean_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    out = torch.zeros((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    out.index_add_(0, batch, x)
    cnt = torch.zeros(n, dtype=x.dtype, device=x.device)
    cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp(min=1).view(-1, *([1] * (x.dim() - 1)))


class Data:
    """Attribute bag with the handful of PyG ``Data`` methods the reference's transforms touch."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to_dict(self):
        return {k: getattr(self, k) for k in self.keys}

    @classmethod
    def from_dict(cls, d):
        o = cls()
        for k, v in d.items():
            setattr(o, k, v)
        return o

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return k in self.__dict__

    def __inc__(self, key, value, *a, **k):
        return 0

    def __cat_dim__(self, key, value, *a, **k):
        return 0


def install():
    """Register the stand-ins and put the reference on sys.path.  Idempotent."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_csmpn_shim", False):
        return
    from . import lift_ref

    tg = types.ModuleType("torch_geometric")
    tg._csmpn_shim = True
    tg.seed = types.ModuleType("torch_geometric.seed")
    tg.seed.seed_everything = lambda s: torch.manual_seed(s)
    tg.seed_everything = tg.seed.seed_everything
    nn_mod = types.ModuleType("torch_geometric.nn")
    nn_mod.MessagePassing = MessagePassing
    nn_mod.global_mean_pool = global_mean_pool
    nn_mod.knn_graph = lift_ref.knn_graph
    data_mod = types.ModuleType("torch_geometric.data")
    data_mod.Data = Data
    data_mod.InMemoryDataset = object
    data_mod.DataLoader = object
    loader_mod = types.ModuleType("torch_geometric.loader")
    loader_mod.DataLoader = object
    tr_mod = types.ModuleType("torch_geometric.transforms")
    tr_mod.BaseTransform = object
    typing_mod = types.ModuleType("torch_geometric.typing")
    typing_mod.Adj = object
    tg.nn, tg.data, tg.loader, tg.transforms, tg.typing = nn_mod, data_mod, loader_mod, tr_mod, typing_mod
    ts = types.ModuleType("torch_scatter")
    ts.scatter = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("torch_scatter stand-in"))
    gd = types.ModuleType("gudhi")
    gd.SimplexTree = lift_ref.SimplexTree
    gd.RipsComplex = lift_ref.RipsComplex
    gd_st = types.ModuleType("gudhi.simplex_tree")
    gd_st.SimplexTree = lift_ref.SimplexTree
    gd.simplex_tree = gd_st
    mods = {
        "torch_geometric": tg, "torch_geometric.seed": tg.seed, "torch_geometric.nn": nn_mod,
        "torch_geometric.data": data_mod, "torch_geometric.loader": loader_mod,
        "torch_geometric.transforms": tr_mod, "torch_geometric.typing": typing_mod,
        "torch_scatter": ts, "gudhi": gd, "gudhi.simplex_tree": gd_st,
    }
    for k, v in mods.items():
        sys.modules.setdefault(k, v)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
