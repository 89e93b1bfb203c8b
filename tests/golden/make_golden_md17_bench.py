#!/usr/bin/env python
"""Golden fixture for the md17 model AT THE BENCHED SIZE: the UNMODIFIED reference model
(csmpn/models/md17_cssmpnn.py: Cl(3,0), num_input=30, num_hidden=32, 5 layers -- csmpn/configs/md17.yaml:29-34) on the
exact 100-complex batch `bench.py`'s train leg uses (bench.make_md17_graphs(100, 2000)), run on the CPU of the build
container through oracle/refshim.py.

    python tests/golden/make_golden_md17_bench.py     (needs /root/reference; writes tests/golden/md17_bench.pt, ~3 MB)

Stored: the state_dict, the training loss, the per-sample losses, the gradient of the loss w.r.t. every parameter and
a checksum of the collated edge_index / x_ind (the batch itself is regenerated from the seed by the test and lifted on
the GPU; the checksum proves both sides saw the same complexes in the same order).
"""
import hashlib
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch

from oracle import lift_ref as L
from oracle import refshim

refshim.install()

N_COMPLEXES, SEED = 100, 2000


def checksum(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def main():
    import bench
    from csmpn.models.md17_cssmpnn import CliffordSharedSimplicialMPNN_md17
    from make_golden_models import collate

    graphs = bench.make_md17_graphs(N_COMPLEXES, SEED, "cpu")
    parts, loc, vel, chg, ys = [], [], [], [], []
    for g in graphs:
        n, F = g.loc.shape[0], g.loc.shape[1]
        parts.append(L.merge_ref(*L.clique_lift_ref(n, g.edge_index)))
        loc.append(g.loc), vel.append(g.vel), ys.append(g.y)
        chg.append(g.charges.reshape(n, 1, 1).repeat(1, F, 1))
    b = collate(parts, dict(loc=loc, vel=vel, charges=chg))
    b["y"] = torch.cat(ys)
    torch.manual_seed(0)
    model = CliffordSharedSimplicialMPNN_md17()
    t0 = time.time()
    loss, out = model(types.SimpleNamespace(**{k: v.clone() for k, v in b.items()}), 0, "train")
    named = list(model.named_parameters())
    grads = torch.autograd.grad(loss, [p for _, p in named], allow_unused=True)
    print(f"reference md17 model fwd+bwd on {b['x_ind'].shape[0]} simplices / {b['edge_index'].shape[1]} pairs: {time.time() - t0:.1f} s")
    fx = dict(
        n_complexes=N_COMPLEXES, seed=SEED, loss=loss.detach(), out={k: v.detach() for k, v in out.items()},
        grads={n: (None if g is None else g.detach()) for (n, _), g in zip(named, grads)},
        state_dict={k: v.detach().clone() for k, v in model.state_dict().items() if "algebra" not in k},
        edge_index_sha256=checksum(b["edge_index"]), x_ind_sha256=checksum(b["x_ind"]),
        n_simplices=int(b["x_ind"].shape[0]), n_pairs=int(b["edge_index"].shape[1]))
    path = os.path.join(HERE, "md17_bench.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


if __name__ == "__main__":
    main()
