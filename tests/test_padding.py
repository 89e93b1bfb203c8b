"""csmpn_b200.data.padding: shape-padded batches (one dummy complex absorbs the difference to a bucket of simplex / pair
counts).  Host logic only -- torch ops on CPU tensors: the structure of the padded batch, in-place reuse, overflow.  The
effect on the model (unchanged loss and gradients, one CUDA graph for a stream of batches) is in tests/test_models.py."""
import pytest
import torch

from oracle import lift_ref as L


def _collated(gen, sizes=(7, 7, 7), frames=2):
    """a collated batch of clique complexes on len(sizes) kNN graphs, built with the oracle lifter (no GPU)"""
    from csmpn_b200.data.modules.simplicial_data import Data

    eis, xinds, types, batch, locs, ys, ptr = [], [], [], [], [], [], [0]
    n_edges = n_tris = 0
    for c, n in enumerate(sizes):
        loc = torch.randn(n, frames, 3, generator=gen) * 1.5
        ei, x_ind, nt = L.merge_ref(*L.clique_lift_ref(n, L.knn_graph(loc[:, 0], 3)))
        N = x_ind.shape[0]
        eis.append(ei + ptr[-1]); xinds.append(x_ind); types.append(nt); batch.append(torch.full((N,), c, dtype=torch.int64))
        locs.append(torch.cat([loc, torch.zeros(N - n, frames, 3)])); ys.append(loc + 0.1)
        n_edges += int((nt == 1).sum()); n_tris += int((nt == 2).sum())
        ptr.append(ptr[-1] + N)
    b = Data(edge_index=torch.cat(eis, 1), x_ind=torch.cat(xinds), node_types=torch.cat(types), batch=torch.cat(batch),
             ptr=torch.tensor(ptr), loc=torch.cat(locs), y=torch.cat(ys), num_graphs=len(sizes))
    b.x_ind_batch = b.node_types_batch = b.batch
    b.x_ind_ptr = b.node_types_ptr = b.ptr
    b.sizes = {"vertices": sum(sizes), "edges": n_edges, "triangles": n_tris, "pairs": int(b.edge_index.shape[1]),
               "simplices": ptr[-1], "complexes": len(sizes)}
    return b


def test_padded_batch_structure():
    from csmpn_b200.data.padding import make_bucket, pad_to_bucket

    gen = torch.Generator().manual_seed(0)
    b = _collated(gen)
    bucket = make_bucket([b.sizes], margin=1.1, multiple=8)
    assert bucket.n_dummy_vertices == 7 and bucket.vertices == 21
    assert bucket.edges > b.sizes["edges"] and bucket.triangles > b.sizes["triangles"] and bucket.pairs > b.sizes["pairs"]
    p = pad_to_bucket(b, bucket)
    n, N = b.sizes["simplices"], bucket.simplices
    assert p.x_ind.shape[0] == p.node_types.shape[0] == p.batch.shape[0] == p.loc.shape[0] == N
    assert p.edge_index.shape == (2, bucket.pairs) and p.y.shape[0] == b.y.shape[0] + 7
    assert [int((p.node_types == d).sum()) for d in range(3)] == list(bucket.counts()) == list(p.pad_counts)
    # the real part is untouched, the dummy complex is last: own graph id, zero features, only self-pairs among its rows
    for k in ("x_ind", "node_types", "batch", "loc"):
        assert torch.equal(getattr(p, k)[:n], getattr(b, k))
    assert torch.equal(p.edge_index[:, : b.sizes["pairs"]], b.edge_index)
    assert bool((p.batch[n:] == 3).all()) and p.ptr.tolist() == b.ptr.tolist() + [N]
    assert p.num_graphs == 4 and p.n_real_graphs == 3
    assert float(p.loc[n:].abs().sum()) == 0.0 and float(p.y[b.y.shape[0]:].abs().sum()) == 0.0
    pad = p.edge_index[:, b.sizes["pairs"]:]
    assert torch.equal(pad[0], pad[1]) and int(pad.min()) >= n and int(pad.max()) < N
    deg = torch.bincount(pad[1] - n, minlength=N - n)
    assert int(deg.max()) - int(deg.min()) <= 1, "padding pairs are dealt round-robin over the dummy simplices"
    # dummy simplices reference vertices of the dummy complex only (local ids < its vertex count)
    assert int(p.x_ind[n:].max()) < bucket.n_dummy_vertices
    assert p.x_ind_batch is p.batch and p.x_ind_ptr is p.ptr


def test_padding_in_place_and_overflow():
    from csmpn_b200.data.padding import BucketOverflow, make_bucket, pad_to_bucket

    gen = torch.Generator().manual_seed(1)
    batches = [_collated(gen) for _ in range(4)]
    assert len({b.sizes["simplices"] for b in batches}) > 1, "the batches should differ in size"
    bucket = make_bucket([b.sizes for b in batches])
    static = pad_to_bucket(batches[0], bucket)
    ptrs = {k: getattr(static, k).data_ptr() for k in ("x_ind", "node_types", "batch", "ptr", "edge_index", "loc", "y")}
    for b in batches[1:]:
        out = pad_to_bucket(b, bucket, out=static)
        fresh = pad_to_bucket(b, bucket)
        assert out is static
        for k, a in ptrs.items():
            assert getattr(static, k).data_ptr() == a, f"{k} was re-allocated"
            assert torch.equal(getattr(static, k), getattr(fresh, k)), k
    small = make_bucket([{**batches[0].sizes, "pairs": batches[0].sizes["pairs"] - 100}], margin=1.0, multiple=1)
    with pytest.raises(BucketOverflow):
        pad_to_bucket(batches[0], small)
    other = dict(batches[0].sizes, vertices=28, complexes=4)
    with pytest.raises(ValueError):
        make_bucket([batches[0].sizes, other])
