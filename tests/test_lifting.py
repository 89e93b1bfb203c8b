"""Simplicial lifting: (1) the CPU oracle (oracle/lift_ref.py) against the fixtures produced by the reference's own
utils.py / simplicial_data.py (tests/golden/lifting.pt, made by tests/golden/make_golden.py), (2) the GPU lifter
(csrc/lift.cu through the C ABI) against both.  Bit-exact: values AND order of every output."""
import itertools

import pytest
import torch

from conftest import load_golden
from oracle import lift_ref as L


@pytest.fixture(scope="module")
def gold():
    return load_golden("lifting.pt")


def oracle_case(name, g):
    if name in ("nba6", "nba11", "rips9_sparse"):
        x, adj = L.rips_lift_ref(g["points"].tolist(), 2, g["dis"])
    elif name.startswith("md17"):
        x, adj = L.clique_lift_ref(g["points"].shape[0], g["knn_edge_index"])
    elif name == "hulls8":
        x, adj = L.hull_faces_lift_ref(8, g["facets"].tolist(), 2)
    else:
        return L.motion_manual_ref(g["base_edge_index"])
    return L.merge_ref(x, adj)


CASES = ["nba6", "nba11", "rips9_sparse", "md17_13", "md17_21", "hulls8", "motion"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(gold, name):
    ei, x_ind, nt = oracle_case(name, gold[name])
    assert torch.equal(ei, gold[name]["edge_index"])
    assert torch.equal(x_ind, gold[name]["x_ind"]) and x_ind.dtype == torch.float32
    assert torch.equal(nt, gold[name]["node_types"])


def test_frozenset_order_known_answers():
    # not sorted once an id >= 8 appears
    assert list(frozenset([6, 8])) == [8, 6]
    assert list(frozenset([7, 8, 9])) == [8, 9, 7]


# ------------------------------------------------------------------------------------------------ GPU
def gpu_lift(name, g, dev):
    from csmpn_b200.data.modules import lifting as G

    if name in ("nba6", "nba11", "rips9_sparse"):
        return G.lift_batch(G.LIFT_RIPS, [g["points"].shape[0]], points=g["points"].to(dev), max_edge_length=g["dis"])
    if name.startswith("md17"):
        ei = g["knn_edge_index"].to(dev)
        return G.lift_batch(G.LIFT_CLIQUE, [g["points"].shape[0]], pairs=ei, pairs_per_complex=[ei.shape[1]])
    if name == "hulls8":
        f = g["facets"].to(dev)
        return G.lift_batch(G.LIFT_FACETS, [8], facets=f, facets_per_complex=[f.shape[0]])
    b = g["base_edge_index"].to(dev)
    return G.lift_batch(G.LIFT_MOTION, [31], pairs=b, pairs_per_complex=[b.shape[1]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_lift_matches_reference_fixture(gold, name):
    dev = torch.device("cuda:0")
    lb = gpu_lift(name, gold[name], dev)
    assert torch.equal(lb.edge_index.cpu(), gold[name]["edge_index"])
    assert torch.equal(lb.x_ind.cpu(), gold[name]["x_ind"]) and lb.x_ind.dtype == torch.float32
    assert torch.equal(lb.node_types.cpu(), gold[name]["node_types"])
    assert int(lb.batch.max()) == 0


def _collate_ref(parts):
    eis, xs, nts, off = [], [], [], 0
    for ei, x_ind, nt in parts:
        eis.append(ei + off), xs.append(x_ind), nts.append(nt)
        off += x_ind.shape[0]
    return torch.cat(eis, 1), torch.cat(xs, 0), torch.cat(nts, 0)


@pytest.mark.gpu
def test_gpu_lift_batched_ragged_rips_vs_oracle():
    """100 ragged point clouds (1..32 vertices, radius from 'no edge' to 'complete') in one launch vs the oracle."""
    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    sizes = [1, 2, 3, 32, 31, 17] + [int(v) for v in torch.randint(4, 14, (58,), generator=gen)]
    for dis in (0.0, 0.9, 1e4):
        pts = [torch.randn(n, 3, generator=gen) for n in sizes]
        if dis == 1e4:   # keep the complete complexes small enough for the pure-Python oracle
            pts = [p[:12] for p in pts]
        nv = [p.shape[0] for p in pts]
        lb = G.lift_batch(G.LIFT_RIPS, nv, points=torch.cat(pts).to(dev), max_edge_length=dis)
        ref = _collate_ref([L.merge_ref(*L.rips_lift_ref(p.tolist(), 2, dis)) if _has_edges(p, dis)
                            else _no_edge_ref(p.shape[0]) for p in pts])
        assert torch.equal(lb.edge_index.cpu(), ref[0]), dis
        assert torch.equal(lb.x_ind.cpu(), ref[1])
        assert torch.equal(lb.node_types.cpu(), ref[2])
        assert lb.node_ptr.cpu().tolist() == [0] + list(itertools.accumulate(int(c[0] + c[1]) + n for c, n in zip(lb.counts.cpu().tolist(), nv)))


def _has_edges(p, dis):
    return p.shape[0] > 1


def _no_edge_ref(n):
    """a single vertex: no adjacency at all (merge_ref cannot concatenate zero blocks)"""
    x_ind = torch.zeros((n, 3))
    x_ind[:, 0] = torch.arange(n).float()
    return torch.zeros((2, 0), dtype=torch.long), x_ind, torch.zeros(n, dtype=torch.long)


@pytest.mark.gpu
def test_gpu_lift_full_32_vertex_complex_properties():
    """maximum size: complete complex on 32 vertices (496 edges, 4960 triangles) -- closed-form counts, every block
    sorted the way the reference emits it, incidences consistent with x_ind."""
    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    pts = torch.randn(32, 2, generator=torch.Generator().manual_seed(0))
    lb = G.lift_batch(G.LIFT_RIPS, [32], points=pts.to(dev), max_edge_length=1e9)
    n, ne, nt = 32, 496, 4960
    assert lb.counts.cpu().tolist() == [[ne, nt]]
    assert lb.x_ind.shape[0] == n + ne + nt
    E = 2 * ne + (n * (n - 1) - ne) + 4 * ne + 12 * nt
    assert lb.edge_index.shape[1] == E
    ei, x = lb.edge_index.cpu(), lb.x_ind.cpu().long()
    # block 1_2: each triangle receives its three faces; the face vertex sets are subsets of the triangle's
    p = 2 * ne + (n * (n - 1) - ne) + 4 * ne + 6 * nt
    blk = ei[:, p:p + 3 * nt]
    assert torch.equal(blk[1], (n + ne + torch.arange(nt)).repeat_interleave(3))
    tri_sets = [set(r.tolist()) for r in x[n + ne:]]
    for col in range(0, 3 * nt, 97):
        e, t = int(blk[0, col]), int(blk[1, col]) - n - ne
        assert set(x[e, :2].tolist()) < tri_sets[t]
    # block 2_1 is block 1_2 reversed
    assert torch.equal(ei[:, p + 3 * nt:p + 6 * nt], blk[[1, 0]])


@pytest.mark.gpu
def test_gpu_lift_clique_and_facets_batched_vs_oracle():
    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    # clique: kNN graphs of 40 random molecules of 5..21 atoms
    parts, pairs, nv, pp = [], [], [], []
    for _ in range(40):
        n = int(torch.randint(5, 22, (1,), generator=gen))
        ei = L.knn_graph(torch.randn(n, 3, generator=gen), 3)
        parts.append(L.merge_ref(*L.clique_lift_ref(n, ei)))
        pairs.append(ei), nv.append(n), pp.append(ei.shape[1])
    lb = G.lift_batch(G.LIFT_CLIQUE, nv, pairs=torch.cat(pairs, 1).to(dev), pairs_per_complex=pp)
    ref = _collate_ref(parts)
    assert torch.equal(lb.edge_index.cpu(), ref[0]) and torch.equal(lb.x_ind.cpu(), ref[1]) and torch.equal(lb.node_types.cpu(), ref[2])
    assert torch.equal(lb.batch.cpu(), torch.repeat_interleave(torch.arange(40), torch.tensor([p[1].shape[0] for p in parts])))
    # facets: convex hulls of 8 points in R^5 (Qhull on the host supplies the facets, as in the reference)
    from scipy.spatial import ConvexHull
    parts, facets, fp = [], [], []
    for _ in range(12):
        f = torch.as_tensor(ConvexHull(torch.randn(8, 5, generator=gen).numpy()).simplices).long()
        parts.append(L.merge_ref(*L.hull_faces_lift_ref(8, f.tolist(), 2)))
        facets.append(f), fp.append(f.shape[0])
    lb = G.lift_batch(G.LIFT_FACETS, [8] * 12, facets=torch.cat(facets).to(dev), facets_per_complex=fp)
    ref = _collate_ref(parts)
    assert torch.equal(lb.edge_index.cpu(), ref[0]) and torch.equal(lb.x_ind.cpu(), ref[1]) and torch.equal(lb.node_types.cpu(), ref[2])


@pytest.mark.gpu
def test_gpu_transform_api_matches_reference_fixture(gold):
    """the reference-facing classes: SimplicialTransform / ManualTransform / rips_lift on single samples"""
    from csmpn_b200.data.modules import utils as U
    from csmpn_b200.data.modules.simplicial_data import Data, ManualTransform, SimplicialTransform

    dev = torch.device("cuda:0")
    g = gold["nba6"]
    pos = torch.zeros(6, 4, 2)
    pos[:, 0] = g["points"]
    d = SimplicialTransform(dim=2, dis=g["dis"], label="nba")(Data(pos=pos.to(dev), vel=torch.randn(6, 4, 2).to(dev),
                                                                   init_pos=g["points"].to(dev), y=torch.zeros(5, 2, 2)))
    assert torch.equal(d.edge_index.cpu(), g["edge_index"]) and torch.equal(d.x_ind.cpu(), g["x_ind"])
    assert d.pos.shape == (41, 4, 2) and torch.equal(d.pos[:6].cpu(), pos) and float(d.pos[6:].abs().max()) == 0.0
    x_dict, adj = U.rips_lift(Data(init_pos=g["points"].to(dev)), 2, g["dis"])
    xr, ar = L.rips_lift_ref(g["points"].tolist(), 2, g["dis"])
    assert all(torch.equal(x_dict[k].cpu(), xr[k]) for k in xr) and set(adj) == set(ar)
    assert all(torch.equal(adj[k].cpu(), ar[k]) for k in ar)
    m = gold["motion"]
    d = ManualTransform()(Data(loc=torch.randn(31, 3).to(dev), vel=torch.randn(31, 3).to(dev), edge_index=m["base_edge_index"].to(dev),
                               y=torch.zeros(31, 3)))
    assert torch.equal(d.edge_index.cpu(), m["edge_index"]) and torch.equal(d.x_ind.cpu(), m["x_ind"])
    assert torch.equal(d.node_types.cpu(), m["node_types"])


# ------------------------------------------------------------------------------------------------ filters + kNN
@pytest.fixture(scope="module")
def gold_filtered():
    return load_golden("lifting_filtered.pt")


FILTERED = ["md17_21_k4_mid", "md17_21_k5_edges_only", "md17_13_k4_tris_only", "md17_21_k6_tight"]


@pytest.mark.parametrize("name", FILTERED)
def test_oracle_filters_match_reference_fixture(gold_filtered, name):
    """edge_th / tri_th of the reference's simplicial_lift (utils.py:181-200): the restatement against fixtures made by
    the reference's own code (tests/golden/make_golden_lift_filtered.py)"""
    g = gold_filtered[name]
    x, adj = L.clique_lift_ref(g["points"].shape[0], g["knn_edge_index"], g["points"], g["edge_th"], g["tri_th"])
    ei, x_ind, nt = L.merge_ref(x, adj)
    assert torch.equal(ei, g["edge_index"]) and torch.equal(x_ind, g["x_ind"]) and torch.equal(nt, g["node_types"])
    assert int((nt == 2).sum()) < g["n_cliques"] or g["tri_th"] >= 1e4          # the area filter removed triangles
    assert int((nt == 1).sum()) <= g["n_graph_edges"]


def test_oracle_knn_graph_matches_fixture(gold_filtered):
    for name in FILTERED:
        g = gold_filtered[name]
        assert torch.equal(L.knn_graph(g["points"], g["knn_k"]), g["knn_edge_index"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", FILTERED)
def test_gpu_lift_filters_and_knn_match_reference_fixture(gold_filtered, name):
    """GPU lifter with the length / area filters, from the kNN pairs (LIFT_CLIQUE) and with the kNN graph built in the
    kernel (LIFT_KNN): bit-exact to the reference fixture"""
    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    g = gold_filtered[name]
    n = g["points"].shape[0]
    ei = g["knn_edge_index"].to(dev)
    a = G.lift_batch(G.LIFT_CLIQUE, [n], pairs=ei, pairs_per_complex=[ei.shape[1]], points=g["points"].to(dev),
                     edge_th=g["edge_th"], tri_th=g["tri_th"])
    b = G.lift_batch(G.LIFT_KNN, [n], points=g["points"].to(dev), knn_k=g["knn_k"], edge_th=g["edge_th"], tri_th=g["tri_th"])
    for lb in (a, b):
        assert torch.equal(lb.edge_index.cpu(), g["edge_index"])
        assert torch.equal(lb.x_ind.cpu(), g["x_ind"]) and torch.equal(lb.node_types.cpu(), g["node_types"])


@pytest.mark.gpu
def test_gpu_knn_lift_batched_equals_pairs_lift(gold):
    """a ragged batch: LIFT_KNN (graph built on the GPU) == LIFT_CLIQUE fed with the oracle's knn_graph pairs, and the
    SimplicialTransform API (knn_k=...) gives the same collated batch as the edge_index path"""
    from csmpn_b200.data.modules import lifting as G
    from csmpn_b200.data.modules.simplicial_data import Data, SimplicialTransform

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    sizes = [21, 13, 5, 32, 2, 21, 9]
    pts = [torch.randn(n, 3, generator=gen) * 1.5 for n in sizes]
    pairs = [L.knn_graph(p, 3) for p in pts]
    a = G.lift_batch(G.LIFT_CLIQUE, sizes, pairs=torch.cat(pairs, 1).to(dev), pairs_per_complex=[p.shape[1] for p in pairs])
    b = G.lift_batch(G.LIFT_KNN, sizes, points=torch.cat(pts).to(dev), knn_k=3)
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.x_ind, b.x_ind) and torch.equal(a.node_types, b.node_types)
    graphs = [Data(loc=p.unsqueeze(1).repeat(1, 2, 1), vel=torch.zeros(p.shape[0], 2, 3), edge_index=e, charges=torch.ones(p.shape[0]),
                   y=torch.zeros(p.shape[0], 2, 3)) for p, e in zip(pts, pairs)]
    t1 = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin").lift(graphs, device=dev)
    t2 = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin", knn_k=3).lift(graphs, device=dev)
    assert torch.equal(t1.edge_index, t2.edge_index) and torch.equal(t1.x_ind, t2.x_ind) and torch.equal(t1.loc, t2.loc)


@pytest.mark.gpu
def test_gpu_transform_api_with_filters(gold_filtered):
    from csmpn_b200.data.modules import utils as U
    from csmpn_b200.data.modules.simplicial_data import Data, SimplicialTransform

    dev = torch.device("cuda:0")
    g = gold_filtered["md17_21_k4_mid"]
    p = g["points"]
    graph = Data(loc=p.unsqueeze(1).repeat(1, 3, 1).to(dev), vel=torch.zeros(21, 3, 3, device=dev), init_pos=p.to(dev),
                 edge_index=g["knn_edge_index"].to(dev), charges=torch.ones(21, device=dev), y=torch.zeros(21, 3, 3, device=dev))
    d = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin", edge_th=g["edge_th"], tri_th=g["tri_th"])(graph)
    assert torch.equal(d.edge_index.cpu(), g["edge_index"]) and torch.equal(d.x_ind.cpu(), g["x_ind"])
    x_dict, adj = U.simplicial_lift(graph, edge_th=g["edge_th"], tri_th=g["tri_th"])
    assert x_dict[1].shape[0] == int((g["node_types"] == 1).sum()) and x_dict[2].shape[0] == int((g["node_types"] == 2).sum())
    a = torch.tensor([[0.0, 0.0, 0.0]]); b = torch.tensor([[1.0, 0.0, 0.0]]); c = torch.tensor([[0.0, 2.0, 0.0]])
    assert float(U.triangle_area(a, b, c)) == 1.0


@pytest.mark.gpu
def test_gpu_lift_33_to_64_vertices_vs_oracle():
    """complexes with more than 32 vertices run the two-word-mask kernels (64-bit adjacency masks, one complex per CTA):
    Rips, clique (from pairs and with the in-kernel kNN graph) and facet lifts of mixed-size batches, bit-exact to the oracle"""
    from scipy.spatial import ConvexHull

    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(77)
    sizes = [33, 64, 5, 40, 32, 57]
    pts = [torch.randn(n, 3, generator=gen) for n in sizes]
    # Rips at a radius that gives a sparse-to-medium complex
    parts = [L.merge_ref(*L.rips_lift_ref(p.tolist(), 2, 0.9)) for p in pts]
    ei, x_ind, nt = _collate_ref(parts)
    lb = G.lift_batch(G.LIFT_RIPS, sizes, points=torch.cat(pts).to(dev), max_edge_length=0.9)
    assert torch.equal(lb.edge_index.cpu(), ei) and torch.equal(lb.x_ind.cpu(), x_ind) and torch.equal(lb.node_types.cpu(), nt)
    # clique complexes of kNN graphs: pairs path and in-kernel kNN
    pairs = [L.knn_graph(p, 4) for p in pts]
    parts = [L.merge_ref(*L.clique_lift_ref(n, e)) for n, e in zip(sizes, pairs)]
    ei, x_ind, nt = _collate_ref(parts)
    a = G.lift_batch(G.LIFT_CLIQUE, sizes, pairs=torch.cat(pairs, 1).to(dev), pairs_per_complex=[e.shape[1] for e in pairs])
    b = G.lift_batch(G.LIFT_KNN, sizes, points=torch.cat(pts).to(dev), knn_k=4)
    for lb in (a, b):
        assert torch.equal(lb.edge_index.cpu(), ei) and torch.equal(lb.x_ind.cpu(), x_ind) and torch.equal(lb.node_types.cpu(), nt)
    # facets: faces of the convex hull triangles of 40 and 64 points in R^3
    fsizes = [40, 64]
    fpts = [torch.randn(n, 3, generator=gen) for n in fsizes]
    facets = [torch.as_tensor(ConvexHull(p.numpy()).simplices).long() for p in fpts]
    parts = [L.merge_ref(*L.hull_faces_lift_ref(n, f.tolist(), 2)) for n, f in zip(fsizes, facets)]
    ei, x_ind, nt = _collate_ref(parts)
    lb = G.lift_batch(G.LIFT_FACETS, fsizes, facets=torch.cat(facets).to(dev), facets_per_complex=[f.shape[0] for f in facets])
    assert torch.equal(lb.edge_index.cpu(), ei) and torch.equal(lb.x_ind.cpu(), x_ind) and torch.equal(lb.node_types.cpu(), nt)
    with pytest.raises(ValueError):
        G.lift_batch(G.LIFT_RIPS, [65], points=torch.randn(65, 3).to(dev), max_edge_length=1.0)


@pytest.mark.gpu
def test_gpu_lift_full_64_vertex_complex_counts():
    """complete complex on 64 vertices: 2 016 edges, 41 664 triangles, closed-form block sizes"""
    from csmpn_b200.data.modules import lifting as G

    dev = torch.device("cuda:0")
    n, ne, nt = 64, 2016, 41664
    lb = G.lift_batch(G.LIFT_RIPS, [n], points=torch.randn(n, 2).to(dev), max_edge_length=1e9)
    assert lb.x_ind.shape[0] == n + ne + nt
    assert lb.edge_index.shape[1] == 6 * ne + 12 * nt + n * (n - 1) - ne
    assert int((lb.node_types == 2).sum()) == nt
    assert int(lb.edge_index.max()) == n + ne + nt - 1 and int(lb.edge_index.min()) == 0
