"""csmpn_b200.data.batcher: the device-side batcher for pre-lifted complexes and the reader of the reference's cached
``(data, slices)`` files -- against a plain restatement of PyG's collate (tests/golden/make_golden_models.py::collate, the
function that produced the model fixtures) on complexes lifted by the oracle.  Runs on the CPU (torch ops only)."""
import os
import sys
import types

import pytest
import torch

from conftest import GOLDEN
from oracle import lift_ref as L

sys.path.insert(0, GOLDEN)


def _samples(gen, n_samples=7, frames=2):
    out = []
    for i in range(n_samples):
        n = [7, 5, 9, 3, 7, 6, 8][i % 7]
        loc = torch.randn(n, frames, 3, generator=gen) * 1.5
        ei, x_ind, nt = L.merge_ref(*L.clique_lift_ref(n, L.knn_graph(loc[:, 0], 3)))
        N = x_ind.shape[0]
        pad = lambda f: torch.cat([f, torch.zeros((N - n,) + tuple(f.shape[1:]))])
        out.append(dict(edge_index=ei, x_ind=x_ind, node_types=nt, loc=pad(loc), vel=pad(torch.randn(n, frames, 3, generator=gen)),
                        y=loc + 0.1))
    return out


def _reference_collate(samples):
    from make_golden_models import collate

    parts = [(s["edge_index"], s["x_ind"], s["node_types"]) for s in samples]
    n_v = [int((s["node_types"] == 0).sum()) for s in samples]
    b = collate(parts, dict(loc=[s["loc"][:n] for s, n in zip(samples, n_v)], vel=[s["vel"][:n] for s, n in zip(samples, n_v)]))
    b["y"] = torch.cat([s["y"] for s in samples])
    return b


@pytest.mark.parametrize("ids", [[0, 1, 2, 3, 4, 5, 6], [3], [6, 0, 2], [1, 1, 5]])
def test_batch_matches_pyg_collate_restatement(ids):
    from csmpn_b200.data.batcher import ProcessedComplexes

    samples = _samples(torch.Generator().manual_seed(3))
    pc = ProcessedComplexes.from_samples(samples, device="cpu")
    assert len(pc) == 7
    got = pc.batch(ids)
    ref = _reference_collate([samples[i] for i in ids])
    for k in ("edge_index", "x_ind", "node_types", "batch", "ptr", "x_ind_batch", "x_ind_ptr", "loc", "vel", "y"):
        assert torch.equal(got[k], ref[k]), k
    assert got.x_ind.dtype == torch.float32 and got.edge_index.dtype == torch.int64
    assert torch.equal(got.node_types_batch, ref["batch"])


def test_reader_takes_pyg_pickles_without_pyg(tmp_path):
    """a file with the structure PyG 2.3.0 writes (Data object whose _store is a GlobalStorage with a _mapping dict, plus the
    slices dict), pickled under the torch_geometric class paths -- read back with no torch_geometric importable"""
    from csmpn_b200.data.batcher import ProcessedComplexes

    samples = _samples(torch.Generator().manual_seed(4), n_samples=4)
    pc = ProcessedComplexes.from_samples(samples, device="cpu")
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage")}
    try:
        mods = {k: types.ModuleType(k) for k in saved}
        GlobalStorage = type("GlobalStorage", (), {"__module__": "torch_geometric.data.storage"})
        DataCls = type("Data", (), {"__module__": "torch_geometric.data.data"})
        mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
        mods["torch_geometric.data.data"].Data = DataCls
        sys.modules.update(mods)
        store = GlobalStorage()
        store._mapping = dict(pc.data)
        data = DataCls()
        data._store = store
        path = os.path.join(tmp_path, "train_data.pt")
        torch.save((data, dict(pc.slices)), path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    back = ProcessedComplexes.load(path, device="cpu")
    a, b = back.batch([2, 0]), pc.batch([2, 0])
    for k in ("edge_index", "x_ind", "node_types", "loc", "y", "ptr"):
        assert torch.equal(a[k], b[k]), k
    pc.save(os.path.join(tmp_path, "plain.pt"))
    again = ProcessedComplexes.load(os.path.join(tmp_path, "plain.pt"), device="cpu")
    assert torch.equal(again.batch([1, 3]).edge_index, pc.batch([1, 3]).edge_index)


@pytest.mark.gpu
def test_device_batcher_equals_one_launch_lift():
    """complexes lifted one by one on the GPU, stored, batched on the device == the same complexes lifted in one launch"""
    from csmpn_b200.data.batcher import ProcessedComplexes
    from csmpn_b200.data.modules.simplicial_data import Data, SimplicialTransform

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(9)
    graphs = []
    for n in (21, 13, 21, 8, 17):
        loc = torch.randn(n, 2, 3, generator=gen) * 1.5
        graphs.append(Data(loc=loc, vel=torch.randn(n, 2, 3, generator=gen), edge_index=L.knn_graph(loc[:, 0], 3),
                           charges=torch.arange(n).float(), y=loc + 0.1))
    tr = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin")
    singles = [tr.lift([g], device=dev) for g in graphs]
    pc = ProcessedComplexes.from_samples(singles, device=dev)
    ids = [4, 0, 2, 1, 3]
    got = pc.batch(ids)
    ref = tr.lift([graphs[i] for i in ids], device=dev)
    for k in ("edge_index", "x_ind", "node_types", "batch", "ptr", "x_ind_ptr", "loc", "vel", "charges", "y"):
        assert torch.equal(got[k], ref[k]), k
