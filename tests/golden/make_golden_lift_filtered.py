#!/usr/bin/env python
"""Golden fixtures for the length / area filters of the reference's ``simplicial_lift`` (csmpn/data/modules/utils.py:181-200):
the UNMODIFIED reference ``SimplicialTransform(label="md17", molecule_type="aspirin", edge_th=..., tri_th=...)`` run over
the gudhi stand-in of oracle/lift_ref.py (networkx is the real one), thresholds chosen so that both filters bite.

    python tests/golden/make_golden_lift_filtered.py      (needs /root/reference; writes tests/golden/lifting_filtered.pt)

Cases with exactly three 3-cliques are skipped on purpose: there the reference's ``torch.cross`` (called without ``dim``)
takes the cross product along the wrong axis (documented in csmpn_b200/data/modules/utils.py::triangle_area).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import torch

from oracle import refshim

refshim.install()


def main():
    from csmpn.data.modules.simplicial_data import SimplicialTransform
    from oracle.lift_ref import knn_graph
    from oracle.refshim import Data

    gen = torch.Generator().manual_seed(23)
    out = {}
    for name, n, k, q_e, q_t in (("md17_21_k4_mid", 21, 4, 0.5, 0.5), ("md17_21_k5_edges_only", 21, 5, 0.3, 2.0),
                                 ("md17_13_k4_tris_only", 13, 4, 2.0, 0.4), ("md17_21_k6_tight", 21, 6, 0.2, 0.25)):
        loc = torch.randn(n, 3, 3, generator=gen) * 1.5
        ei = knn_graph(loc[:, 0], k)
        p = loc[:, 0]
        und = {(min(a, b), max(a, b)) for a, b in ei.t().tolist()}
        lens = torch.tensor(sorted(float(torch.norm(p[a] - p[b])) for a, b in und))
        nbr = [set() for _ in range(n)]
        for a, b in und:
            nbr[a].add(b), nbr[b].add(a)
        areas = sorted(float(0.5 * torch.linalg.norm(torch.linalg.cross(p[b] - p[a], p[c] - p[a])))
                       for a, b in und for c in nbr[a] & nbr[b] if c > b)
        assert len(areas) != 3 and len(areas) > 0, (name, len(areas))
        # thresholds half-way between two sorted values: no ties with the fp32 evaluation on either side
        pick = lambda xs, q: (1e4 if q > 1 else 0.5 * (float(xs[int(q * (len(xs) - 1))]) + float(xs[int(q * (len(xs) - 1)) + 1])))
        edge_th, tri_th = pick(lens, q_e), pick(torch.tensor(areas), q_t)
        g = Data(loc=loc.clone(), vel=torch.randn(n, 3, 3, generator=gen), init_pos=p.clone(), edge_index=ei,
                 y=torch.zeros(n, 3, 3), charges=torch.arange(n).float())
        tr = SimplicialTransform(dim=2, dis=2.5, label="md17", molecule_type="aspirin", edge_th=edge_th, tri_th=tri_th)
        d = tr(g)
        out[name] = dict(points=p.clone(), knn_edge_index=ei.clone(), knn_k=k, edge_th=edge_th, tri_th=tri_th,
                         edge_index=d.edge_index.clone(), x_ind=d.x_ind.clone(), node_types=d.node_types.clone(),
                         n_graph_edges=len(und), n_cliques=len(areas))
        print(name, "graph edges", len(und), "3-cliques", len(areas), "-> simplices", tuple(d.x_ind.shape), "pairs", tuple(d.edge_index.shape),
              "kept edges", int((d.node_types == 1).sum()), "kept triangles", int((d.node_types == 2).sum()))
    path = os.path.join(HERE, "lifting_filtered.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
