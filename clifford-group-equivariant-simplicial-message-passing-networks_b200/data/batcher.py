"""Device-side batcher for PRE-LIFTED complexes and a reader for the reference's cached datasets.

The reference lifts every sample once (``pre_transform=SimplicialTransform``), stores all of them in ONE file
``processed/<partition>_data.pt`` = ``(data, slices)`` (PyG 2.3.0 ``InMemoryDataset.collate``: per attribute one tensor
concatenated over samples along ``__cat_dim__`` -- dim 1 for ``edge_index`` / ``adj_*``, dim 0 otherwise, WITHOUT index
increments -- plus ``slices[key]`` = the sample boundaries; csmpn/data/md17.py:88,105-106, simplicial_data.py:9-25), and
then collates mini-batches on the CPU with ``DataLoader(follow_batch=["node_types", "x_ind"])`` (md17.py:136-150):
``edge_index`` shifted by the running simplex count, ``x_ind`` (float, LOCAL vertex ids) left alone,
``batch`` / ``ptr`` / ``x_ind_batch`` / ``x_ind_ptr`` / ``node_types_batch`` / ``node_types_ptr`` added.

``ProcessedComplexes`` keeps that storage layout ON THE DEVICE and builds a mini-batch for any list of sample ids with a
handful of device-side index operations (ragged gather by ``repeat_interleave`` / ``arange``, no Python loop over
samples, no per-sample host objects): lifted complexes never round-trip through Python ``Data`` objects again.

PyG is not installable in the build container, so the reader cannot be checked against a file written by PyG itself
("parity unpinned" for the file format): it follows the pinned version's documented layout and unpickles PyG's
``Data`` / ``GlobalStorage`` classes as plain attribute bags.
"""
from __future__ import annotations

import io
import pickle

import torch

from .modules.simplicial_data import Data

_CAT_DIM1 = ("edge_index", "adj")        # SimplicialComplexData.__cat_dim__ (simplicial_data.py:20-24)
_FOLLOW = ("node_types", "x_ind")        # DataLoader(follow_batch=...) of the reference (md17.py:136, nba.py:112)


def _cat_dim(key: str) -> int:
    return 1 if any(k in key for k in _CAT_DIM1) else 0


class _Bag:
    """stand-in for any torch_geometric class found in a pickle: state lands in __dict__"""

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            self.__dict__.update(state[1])


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("torch_geometric"):
            return type(name, (_Bag,), {})
        return super().find_class(module, name)


class _PickleModule:
    """pickle_module for torch.load that tolerates PyG classes without PyG"""
    __name__ = "csmpn_b200_pyg_tolerant_pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load())
    dump, dumps, Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    HIGHEST_PROTOCOL, DEFAULT_PROTOCOL = pickle.HIGHEST_PROTOCOL, pickle.DEFAULT_PROTOCOL


def _attributes(obj) -> dict:
    """the tensor attributes of a PyG Data object (2.x: Data._store is a GlobalStorage whose _mapping holds them), of an
    attribute bag, or of a dict"""
    if isinstance(obj, dict):
        return {k: v for k, v in obj.items() if torch.is_tensor(v)}
    d = getattr(obj, "__dict__", {})
    store = d.get("_store")
    if store is not None:
        m = getattr(store, "__dict__", {}).get("_mapping")
        if isinstance(m, dict):
            return {k: v for k, v in m.items() if torch.is_tensor(v)}
    return {k: v for k, v in d.items() if torch.is_tensor(v) and not k.startswith("_")}


class ProcessedComplexes:
    """All samples of a dataset in the reference's ``(data, slices)`` layout, resident on ``device``."""

    def __init__(self, data: dict, slices: dict, device="cuda"):
        self.device = torch.device(device)
        keys = [k for k in data if k in slices]
        if "x_ind" not in keys or "edge_index" not in keys:
            raise ValueError("a processed simplicial dataset holds at least edge_index and x_ind")
        self.data = {k: data[k].to(self.device) for k in keys}
        self.slices = {k: torch.as_tensor(slices[k], dtype=torch.int64).to(self.device) for k in keys}
        self.n_samples = int(self.slices["x_ind"].numel()) - 1
        self._sizes = {k: (s[1:] - s[:-1]) for k, s in self.slices.items()}

    # ---- constructors ----------------------------------------------------------------------------------------
    @classmethod
    def from_samples(cls, samples, device="cuda"):
        """samples: per-sample lifted complexes (objects or dicts with edge_index, x_ind, node_types, features...),
        i.e. what ``SimplicialTransform.__call__`` returns -- concatenated exactly like ``InMemoryDataset.collate``."""
        attrs = [_attributes(s) for s in samples]
        keys = [k for k in attrs[0] if all(k in a for a in attrs) and attrs[0][k].dim() > 0]
        data, slices = {}, {}
        for k in keys:
            dim = _cat_dim(k)
            parts = [a[k] for a in attrs]
            data[k] = torch.cat(parts, dim=dim)
            sizes = torch.tensor([0] + [p.shape[dim] for p in parts])
            slices[k] = torch.cumsum(sizes, 0)
        return cls(data, slices, device)

    @classmethod
    def load(cls, path, device="cuda"):
        """the reference's ``processed/<partition>_data.pt`` (torch.save((data, slices), path), md17.py:105-106)"""
        obj = torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)
        if not (isinstance(obj, (tuple, list)) and len(obj) >= 2):
            raise ValueError(f"{path}: expected the (data, slices) pair PyG's InMemoryDataset stores")
        return cls(_attributes(obj[0]), dict(obj[1]), device)

    def save(self, path):
        """write the same two-element structure (plain dicts) -- readable by ``load``"""
        torch.save(({k: v.cpu() for k, v in self.data.items()}, {k: v.cpu() for k, v in self.slices.items()}), path)

    def __len__(self):
        return self.n_samples

    # ---- the batcher -----------------------------------------------------------------------------------------
    def batch(self, ids) -> Data:
        """collated mini-batch of samples ``ids`` (list / tensor), identical to what the reference's DataLoader builds:
        every attribute gathered sample after sample (ragged gather on the device), ``edge_index`` / ``adj_*`` shifted by the
        running simplex count, ``batch`` / ``ptr`` and the ``follow_batch`` vectors added"""
        dev = self.device
        ids = torch.as_tensor(ids, dtype=torch.int64, device=dev)
        n = int(ids.numel())
        out = Data(num_graphs=n)
        keys = list(self.data)
        sizes = {k: self._sizes[k][ids] for k in keys}
        totals = torch.stack([sizes[k].sum() for k in keys]).cpu().tolist() if n else [0] * len(keys)  # the one host sync
        size_n = sizes["x_ind"]
        node_first = torch.cumsum(size_n, 0) - size_n
        for key, total in zip(keys, totals):
            size, dim = sizes[key], _cat_dim(key)
            first = torch.cumsum(size, 0) - size
            pos = torch.repeat_interleave(self.slices[key][ids] - first, size, output_size=total) + torch.arange(total, device=dev)
            v = self.data[key].index_select(dim, pos)
            if _cat_dim(key) == 1:  # PyG increments edge_index / adj_* by the node count of the preceding samples
                v = v + torch.repeat_interleave(node_first, size, output_size=total).unsqueeze(0)
            out[key] = v
        total_n = totals[keys.index("x_ind")]
        seg = torch.repeat_interleave(torch.arange(n, device=dev), size_n, output_size=total_n)
        ptr = torch.cat([node_first, (node_first[-1:] + size_n[-1:])]) if n else torch.zeros(1, dtype=torch.int64, device=dev)
        out.batch, out.ptr = seg, ptr
        for k in _FOLLOW:
            if k in self.data:
                out[k + "_batch"], out[k + "_ptr"] = seg, ptr
        return out
