#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            (needs /root/reference; writes tests/golden/*.pt)

Layers: the reference's own csmpn.algebra / csmpn.models.cegnn_utils classes (PyG's MessagePassing replaced by
the stand-in in oracle/refshim.py) are fed seeded inputs and parameters; inputs, parameters, outputs and
autograd gradients are stored.  Lifting: the reference's csmpn/data/modules/{utils,simplicial_data}.py run over the
gudhi stand-in of oracle/lift_ref.py.  The GPU box has no /root/reference: tests only read the .pt files.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import layers_ref as R
from oracle import refshim

refshim.install()
from csmpn.algebra.cliffordalgebra import CliffordAlgebra  # noqa: E402  (the reference)
from csmpn.models import cegnn_utils as ref  # noqa: E402


def load_params(module, params, prefix=""):
    sd = {prefix + k if False else k: v for k, v in params.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".algebra." in m or m.startswith("algebra.") for m in missing), missing


def grads_of(out, cot, tensors):
    g = torch.autograd.grad(out, tensors, cot, allow_unused=True)
    return [None if x is None else x.detach().clone() for x in g]


def layer_fixture(metric, C, rows, seed):
    gen = torch.Generator().manual_seed(seed)
    alg = CliffordAlgebra(metric)
    ra = R.RefAlgebra(metric)
    B, G = alg.n_blades, alg.n_subspaces
    fx = {"metric": list(map(float, metric)), "C": C, "rows": rows}

    def rn(*s, std=1.0):
        return torch.randn(*s, generator=gen) * std

    # ---- algebra tables
    fx["cayley"] = alg.cayley.clone()
    fx["subspaces"] = alg.subspaces.clone()
    fx["bbo_grades"] = alg.bbo_grades.clone()
    fx["paths"] = alg.geometric_product_paths.clone()
    fx["index_to_bitmap"] = alg.bbo.index_to_bitmap.clone()

    # ---- geometric product
    a = rn(rows, C, B).requires_grad_()
    b = rn(rows, C, B).requires_grad_()
    out = alg.geometric_product(a, b)
    cot = rn(rows, C, B)
    ga, gb = grads_of(out, cot, [a, b])
    fx["gp"] = dict(a=a.detach(), b=b.detach(), out=out.detach(), cot=cot, ga=ga, gb=gb)
    # norms / qs
    fx["qs"] = torch.cat(alg.qs(a.detach()), dim=-1)
    fx["norms"] = torch.cat(alg.norms(a.detach()), dim=-1)
    fx["norm_all"] = alg.norm(a.detach())

    # ---- MVLinear
    for sub in (True, False):
        cin, cout = C + 3, C
        m = ref.MVLinear(alg, cin, cout, subspaces=sub)
        with torch.no_grad():
            m.bias.copy_(rn(1, cout, 1, std=0.2))
        x = rn(rows, cin, B).requires_grad_()
        y = m(x)
        cot = rn(rows, cout, B)
        gx, gw, gbias = grads_of(y, cot, [x, m.weight, m.bias])
        fx[f"mvlinear_{int(sub)}"] = dict(x=x.detach(), weight=m.weight.detach().clone(), bias=m.bias.detach().clone(),
                                          y=y.detach(), cot=cot, gx=gx, gw=gw, gb=gbias)

    # ---- MVSiLU
    m = ref.MVSiLU(alg, C)
    with torch.no_grad():
        m.a.add_(rn(1, C, G, std=0.2))
        m.b.add_(rn(1, C, G, std=0.2))
    x = rn(rows, C, B).requires_grad_()
    y = m(x)
    cot = rn(rows, C, B)
    gx, ga_, gb_ = grads_of(y, cot, [x, m.a, m.b])
    fx["mvsilu"] = dict(x=x.detach(), a=m.a.detach().clone(), b=m.b.detach().clone(), y=y.detach(), cot=cot, gx=gx, ga=ga_, gb=gb_)

    # ---- NormalizationLayer
    m = ref.NormalizationLayer(alg, C)
    with torch.no_grad():
        m.a.add_(rn(C, G, std=0.5))
    x = rn(rows, C, B).requires_grad_()
    y = m(x)
    cot = rn(rows, C, B)
    gx, ga_ = grads_of(y, cot, [x, m.a])
    fx["mvnorm"] = dict(x=x.detach(), a=m.a.detach().clone(), y=y.detach(), cot=cot, gx=gx, ga=ga_)

    # ---- MVLayerNorm
    m = ref.MVLayerNorm(alg, C)
    with torch.no_grad():
        m.a.add_(rn(1, C, std=0.2))
    x = rn(rows, C, B).requires_grad_()
    y = m(x)
    cot = rn(rows, C, B)
    gx, ga_ = grads_of(y, cot, [x, m.a])
    fx["mvlayernorm"] = dict(x=x.detach(), a=m.a.detach().clone(), y=y.detach(), cot=cot, gx=gx, ga=ga_)

    # ---- SteerableGeometricProductLayer
    m = ref.SteerableGeometricProductLayer(alg, C)
    with torch.no_grad():
        m.normalization.a.add_(rn(C, G, std=0.3))
        m.linear_left.bias.copy_(rn(1, C, 1, std=0.2))
    x = rn(rows, C, B).requires_grad_()
    y = m(x)
    cot = rn(rows, C, B)
    names = [n for n, _ in m.named_parameters()]
    g = grads_of(y, cot, [x] + [p for _, p in m.named_parameters()])
    fx["sgp"] = dict(x=x.detach(), y=y.detach(), cot=cot, gx=g[0],
                     params={n: p.detach().clone() for n, p in m.named_parameters()},
                     grads={n: gi for n, gi in zip(names, g[1:])})

    # ---- CEMLP (2 blocks) with perturbed parameters
    cin = C + 2
    params = R.init_cemlp_params(ra, cin, C, C, 2, gen)
    m = ref.CEMLP(alg, cin, C, C, n_layers=2)
    load_params(m, params)
    x = rn(rows, cin, B).requires_grad_()
    y = m(x)
    cot = rn(rows, C, B)
    plist = [(n, p) for n, p in m.named_parameters()]
    g = grads_of(y, cot, [x] + [p for _, p in plist])
    fx["cemlp"] = dict(x=x.detach(), y=y.detach(), cot=cot, gx=g[0], params=params,
                       grads={n: gi for (n, _), gi in zip(plist, g[1:])})

    # ---- EGCL, both aggregations, random multigraph incl. an isolated receiver and duplicate pairs
    T = 3
    N, E = 14, 57
    for aggr in ("sum", "mean"):
        params = R.init_egcl_params(ra, C, T, gen)
        m = ref.EGCL(alg, C, C, C, edge_attr_features=2 * T, node_attr_features=T, aggr=aggr)
        load_params(m, params)
        h = rn(N, C, B).requires_grad_()
        ei = torch.randint(0, N - 1, (2, E), generator=gen)  # node N-1 never receives nor sends
        ei[:, -1] = ei[:, 0]                                 # one duplicated pair
        node_attr = torch.zeros(N, T, B)
        node_attr[..., 0] = rn(N, T)
        node_attr.requires_grad_()
        edge_attr = torch.cat([node_attr[ei[0]], node_attr[ei[1]]], dim=1).detach().requires_grad_()
        y = m(h, ei, edge_attr=edge_attr, node_attr=node_attr)
        cot = rn(N, C, B)
        plist = [(n, p) for n, p in m.named_parameters()]
        g = grads_of(y, cot, [h, edge_attr, node_attr] + [p for _, p in plist])
        fx[f"egcl_{aggr}"] = dict(h=h.detach(), edge_index=ei, edge_attr=edge_attr.detach(), node_attr=node_attr.detach(),
                                  y=y.detach(), cot=cot, gh=g[0], gedge_attr=g[1], gnode_attr=g[2], params=params,
                                  grads={n: gi for (n, _), gi in zip(plist, g[2 + 1:])})
    return fx


def lifting_fixtures():
    from csmpn.data.modules import utils as U  # the reference, over the gudhi stand-in
    from csmpn.data.modules.simplicial_data import ManualTransform, SimplicialTransform
    from oracle.lift_ref import knn_graph
    from oracle.refshim import Data

    gen = torch.Generator().manual_seed(7)
    out = {}

    def run_transform(tr, g):
        d = tr(g)
        return dict(edge_index=d.edge_index.clone(), x_ind=d.x_ind.clone(), node_types=d.node_types.clone())

    # NBA-shaped: 6 vertices (5 players + reference point), Rips with a huge radius -> full complex; also 11 and a sparse radius
    for name, n, dis in (("nba6", 6, 1e4), ("nba11", 11, 1e4), ("rips9_sparse", 9, 1.2)):
        pos = torch.rand(n, 4, 2, generator=gen) * torch.tensor([47.0, 50.0]) if dis > 100 else torch.randn(n, 4, 2, generator=gen)
        g = Data(pos=pos.clone(), vel=torch.randn(n, 4, 2, generator=gen), init_pos=pos[:, 0].clone(),
                 edge_index=knn_graph(pos[:, 0], 10000), y=torch.zeros(n - 1, 2, 2))
        tr = SimplicialTransform(dim=2, dis=dis, label="nba")
        out[name] = dict(points=pos[:, 0].clone(), dis=dis, **run_transform(tr, g))

    # MD17-shaped: kNN(k=3) clique complex, 13 and 21 atoms
    for name, n in (("md17_13", 13), ("md17_21", 21)):
        loc = torch.randn(n, 3, 3, generator=gen) * 1.5
        ei = knn_graph(loc[:, 0], 3)
        g = Data(loc=loc.clone(), vel=torch.randn(n, 3, 3, generator=gen), init_pos=loc[:, 0].clone(), edge_index=ei,
                 y=torch.zeros(n, 3, 3), charges=torch.arange(n).float())
        tr = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin")
        out[name] = dict(points=loc[:, 0].clone(), knn_edge_index=ei.clone(), **run_transform(tr, g))

    # hulls: 8 points in R^5, all <=2-faces of the Qhull facets
    from scipy.spatial import ConvexHull
    pts = torch.randn(8, 5, generator=gen)
    hull = ConvexHull(pts.numpy())
    full = torch.tensor([[i, j] for i in range(8) for j in range(8) if i != j]).T
    g = Data(input=pts.clone(), target=torch.tensor(float(hull.volume)), edge_index=full, num_nodes=8, y=None)
    tr = SimplicialTransform(dim=2, label="hulls")
    out["hulls8"] = dict(points=pts.clone(), facets=torch.tensor(np.asarray(hull.simplices)).long(), **run_transform(tr, g))

    # motion: the literal ManualTransform complex on a synthetic 31-joint skeleton (60 1-hop + 70 2-hop directed pairs)
    base = synthetic_skeleton_pairs()
    g = Data(loc=torch.randn(31, 3, generator=gen), vel=torch.randn(31, 3, generator=gen), edge_index=base, y=torch.zeros(31, 3))
    d = ManualTransform()(g)
    out["motion"] = dict(base_edge_index=base.clone(), edge_index=d.edge_index.clone(), x_ind=d.x_ind.clone(),
                         node_types=d.node_types.clone())
    return out


def synthetic_skeleton_pairs():
    """A 31-node tree with sum_v C(deg v, 2) = 35 two-hop pairs: 60 + 70 = 130 directed 0-0 pairs, the size of
    the CMU skeleton graph the reference uses."""
    parent = {i: i - 1 for i in range(1, 25)}          # a 25-node path has 23 two-hop pairs
    for child, p in zip(range(25, 31), (3, 7, 11, 15, 19, 22)):   # six leaves on distinct inner joints: +2 each
        parent[child] = p
    und = [(c, p) for c, p in parent.items()]
    adj = {i: set() for i in range(31)}
    for a, b in und:
        adj[a].add(b), adj[b].add(a)
    two = set()
    for v in range(31):
        for a in adj[v]:
            for b in adj[v]:
                if a != b and b not in adj[a]:
                    two.add((a, b))
    rows, cols = [], []
    for i in range(31):
        for j in range(31):
            if i != j and (j in adj[i] or (i, j) in two):
                rows.append(i), cols.append(j)
    return torch.tensor([rows, cols])


def main():
    torch.manual_seed(0)
    cfgs = [("cl2", (1, 1), 6, 5, 11), ("cl3", (1, 1, 1), 8, 5, 12), ("cl5", (1, 1, 1, 1, 1), 4, 3, 13),
            ("cl1m1p1", (1, -1, 1), 4, 4, 14)]
    for name, metric, C, rows, seed in cfgs:
        fx = layer_fixture(metric, C, rows, seed)
        path = os.path.join(HERE, f"layers_{name}.pt")
        torch.save(fx, path)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    lf = lifting_fixtures()
    path = os.path.join(HERE, "lifting.pt")
    torch.save(lf, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", {k: tuple(v["edge_index"].shape) for k, v in lf.items()})


if __name__ == "__main__":
    main()
