#!/usr/bin/env python
"""Golden fixtures for the four task models: the UNMODIFIED reference models (csmpn/models/*_cssmpnn.py, imported from
/root/reference through oracle/refshim.py) run on small seeded batches in the build container.

    python tests/golden/make_golden_models.py        (needs /root/reference; writes tests/golden/models.pt)

Stored per model: constructor kwargs, the collated batch (lifted with oracle/lift_ref.py), the state_dict, the
training loss, the per-sample losses and the gradient of the loss w.r.t. every parameter.  Small widths / few layers
keep the file small; the layer stack is the same code at any width.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import torch

from oracle import lift_ref as L
from oracle import refshim

refshim.install()


def collate(parts, feats):
    """parts: [(edge_index, x_ind, node_types)] per complex; feats: dict name -> list of per-complex vertex features."""
    eis, xs, nts, batch, ptr, off = [], [], [], [], [0], 0
    for c, (ei, x_ind, nt) in enumerate(parts):
        eis.append(ei + off), xs.append(x_ind), nts.append(nt)
        batch.append(torch.full((x_ind.shape[0],), c, dtype=torch.long))
        off += x_ind.shape[0]
        ptr.append(off)
    b = dict(edge_index=torch.cat(eis, 1), x_ind=torch.cat(xs), node_types=torch.cat(nts), batch=torch.cat(batch),
             ptr=torch.tensor(ptr))
    b["x_ind_batch"], b["x_ind_ptr"] = b["batch"].clone(), b["ptr"].clone()
    for name, per in feats.items():
        rows = []
        for (ei, x_ind, nt), f in zip(parts, per):
            pad = torch.zeros((x_ind.shape[0],) + tuple(f.shape[1:]))
            pad[: f.shape[0]] = f
            rows.append(pad)
        b[name] = torch.cat(rows)
    return b


def run(model, batch):
    g = types.SimpleNamespace(**{k: v.clone() for k, v in batch.items()})
    loss, out = model(g, 0, "train")
    names = [n for n, p in model.named_parameters()]
    grads = torch.autograd.grad(loss, [p for _, p in model.named_parameters()], allow_unused=True)
    return dict(loss=loss.detach(), out={k: v.detach() for k, v in out.items()},
                grads={n: (None if gr is None else gr.detach()) for n, gr in zip(names, grads)},
                state_dict={k: v.detach().clone() for k, v in model.state_dict().items() if "algebra" not in k})


def main():
    from csmpn.models.hulls_cssmpnn import HullsCliffordSharedSimplicialMPNN
    from csmpn.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17
    from csmpn.models.motion_cssmpnn import MotionCliffordSharedSimplicialMPNN
    from csmpn.models.nba_cssmpnn import NBACliffordSharedSimplicialMPNN
    from make_golden import synthetic_skeleton_pairs

    gen = torch.Generator().manual_seed(11)
    rn = lambda *s: torch.randn(*s, generator=gen)
    fx = {}

    # ---- md17-shaped: 3 molecules of 7 atoms, 2 frames
    F, n = 2, 7
    parts, loc, vel, chg, ys = [], [], [], [], []
    for _ in range(3):
        l = rn(n, F, 3) * 1.5
        parts.append(L.merge_ref(*L.clique_lift_ref(n, L.knn_graph(l[:, 0], 3))))
        loc.append(l), vel.append(rn(n, F, 3)), chg.append(torch.randint(1, 9, (n, 1, 1), generator=gen).float().repeat(1, F, 1))
        ys.append(l + 0.1 * rn(n, F, 3))
    b = collate(parts, dict(loc=loc, vel=vel, charges=chg))
    b["y"] = torch.cat(ys)
    kw = dict(num_input=3 * F, num_hidden=8, num_out=F, num_layers=2)
    torch.manual_seed(1)
    fx["md17"] = dict(kwargs=kw, batch=b, **run(CliffordSharedSimplicialMPNN_md17(**kw), b))

    # ---- motion-shaped: 3 skeletons on the fixed template
    base = synthetic_skeleton_pairs()
    parts, pos, vel, ys = [], [], [], []
    for _ in range(3):
        parts.append(L.motion_manual_ref(base))
        p = rn(31, 3)
        pos.append(p), vel.append(0.1 * rn(31, 3)), ys.append(p + 0.1 * rn(31, 3))
    b = collate(parts, dict(pos=pos, vel=vel))
    b["y"] = torch.cat(ys)
    kw = dict(num_hidden=8, num_layers=2)
    torch.manual_seed(2)
    fx["motion"] = dict(kwargs=kw, batch=b, **run(MotionCliffordSharedSimplicialMPNN(**kw), b))

    # ---- nba-shaped: 3 complexes of 6 vertices (5 players + reference point), 2 frames, Cl(2,0)
    F, n = 2, 6
    parts, pos, vel, ys = [], [], [], []
    for _ in range(3):
        p = torch.rand(n, F, 2, generator=gen) * torch.tensor([47.0, 50.0])
        parts.append(L.merge_ref(*L.rips_lift_ref(p[:, 0].tolist(), 2, 1e4)))
        pos.append(p), vel.append(rn(n, F, 2)), ys.append(rn(n - 1, 4 * F, 2))
    b = collate(parts, dict(pos=pos, vel=vel))
    b["y"] = torch.cat(ys)
    kw = dict(num_input=2 * F, num_hidden=8, num_out=4 * F, num_layers=2)
    torch.manual_seed(3)
    fx["nba"] = dict(kwargs=kw, batch=b, **run(NBACliffordSharedSimplicialMPNN(**kw), b))

    # ---- hulls-shaped: 2 clouds of 8 points in R^5, Cl(5,0)
    from scipy.spatial import ConvexHull
    parts, inp, tgt = [], [], []
    for _ in range(2):
        p = rn(8, 5)
        hull = ConvexHull(p.numpy())
        parts.append(L.merge_ref(*L.hull_faces_lift_ref(8, torch.as_tensor(hull.simplices).tolist(), 2)))
        inp.append(p), tgt.append(float(hull.volume))
    b = collate(parts, dict(input=inp))
    b["target"] = torch.tensor(tgt)
    kw = dict(hidden_features=4, num_layers=2)
    torch.manual_seed(4)
    fx["hulls"] = dict(kwargs=kw, batch=b, **run(HullsCliffordSharedSimplicialMPNN(**kw), b))

    path = os.path.join(HERE, "models.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB",
          {k: (float(v["loss"]), tuple(v["batch"]["edge_index"].shape)) for k, v in fx.items()})


if __name__ == "__main__":
    main()
