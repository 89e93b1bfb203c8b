"""NBA player-trajectory model with the reference's interface (csmpn/models/nba_cssmpnn.py:12-193); Cl(2,0)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..algebra.cliffordalgebra import CliffordAlgebra
from ._shared import Loss, MetricCollection, SharedSimplicialBase
from .cegnn_utils import CEMLP, EGCL, MVLinear


class NBACliffordSharedSimplicialMPNN(SharedSimplicialBase):
    def __init__(self, max_dim: int = 2, num_input: int = 20, num_hidden: int = 40, num_out: int = 40, num_layers: int = 4,
                 stats=None, condition=True, vertices_per_complex: int = 6) -> None:
        super().__init__()
        alg = self.algebra = CliffordAlgebra((1, 1))
        self.max_dim, self.condition = max_dim, condition
        self.num_input, self.num_hidden, self.num_out = num_input, num_hidden, num_out
        T = self.num_node_type = max_dim + 1 if condition else 0
        # the reference hard-codes six vertices per complex (five players + a reference point, nba_cssmpnn.py:182);
        # the synthetic "10 players + ball" shape of BASELINE.json needs 11
        self.vertices_per_complex = vertices_per_complex
        self.feature_embedding = MVLinear(alg, num_input + T, num_hidden, subspaces=False)
        self.cl_feature_embedding = nn.Sequential(
            MVLinear(alg, num_input, num_input, subspaces=False),
            CEMLP(alg, 2 * num_input, num_hidden, num_input, n_layers=1, normalization_init=0),
            nn.Sequential(CEMLP(alg, 3 * num_input, num_hidden, num_hidden, n_layers=1, normalization_init=0),
                          CEMLP(alg, num_hidden, num_hidden, num_input, n_layers=1, normalization_init=0)))
        self.sim_type_embedding = nn.Embedding(max_dim + 1, max_dim + 1)
        self.stats = stats
        self.layers = nn.Sequential(*[
            EGCL(alg, num_hidden, num_hidden, num_hidden, edge_attr_features=2 * T, node_attr_features=T, aggr="sum",
                 normalization_init=0) for _ in range(num_layers)])
        self.projection = MVLinear(alg, num_hidden, num_out)
        self.train_metrics, self.val_metrics, self.test_metrics = (self._setup_metrics() for _ in range(3))
        self.loss_func = nn.MSELoss(reduction="none")

    def _setup_metrics(self):
        return MetricCollection({"loss": Loss(), "ade_loss": Loss(), "fde_loss": Loss()})

    vertex_feature_types = 2

    def vertex_features(self, graph, verts):
        rows = verts.shape[0]
        pos = self.grade1(graph.pos[verts].reshape(rows, -1, 2))
        vel = self.grade1(graph.vel[verts].reshape(rows, -1, 2))
        return torch.cat((pos, vel), dim=1)

    def featurization(self, x, node_attr):
        return self.feature_embedding(torch.cat((x, node_attr), dim=1))

    def forward(self, graph, step, mode):
        self.begin_forward(graph)
        batch_size = graph.ptr.shape[0] - 1
        n_real = int(getattr(graph, "n_real_graphs", batch_size))  # shape-padded batches end with one dummy complex
        num_frames = graph.pos.shape[1]
        d = self.algebra.dim
        node_attr, edge_attr = self.embed_simplex_types(graph)
        x = self.embed_simplicial_complex(graph, out_channels=self.num_input)
        x = self.featurization(x, node_attr)
        x = self.run_layers(x, graph, edge_attr, node_attr)
        out = self.projection(x[self.simplex_rows(graph)[0]])
        loc_pred = out[..., 1:3].reshape(batch_size, self.vertices_per_complex, num_frames * 4, -1)[:, :-1, ...]
        targets = graph.y
        if n_real != batch_size:
            # the dummy complex leaves BEFORE the losses: its prediction is exactly zero (zero vectors in, zero vectors out,
            # by equivariance) and sqrt has no gradient there
            loc_pred = loc_pred[:n_real]
            targets = targets[: n_real * (self.vertices_per_complex - 1)]
            batch_size = n_real
        loc_pred = loc_pred.reshape(-1, self.num_out, d)
        ade_loss = torch.sqrt(F.mse_loss(loc_pred.reshape(-1, d), targets.reshape(-1, d), reduction="none").sum(dim=-1)
                              ).reshape(batch_size, -1, num_frames).mean(dim=-1).mean(dim=-1)
        fde_loss = torch.sqrt(F.mse_loss(loc_pred[:, -1, :], targets[:, -1, :], reduction="none").sum(dim=-1)
                              ).reshape(batch_size, -1).mean(dim=-1)
        loss = ade_loss
        return loss.mean(), {"loss": loss, "ade_loss": ade_loss, "fde_loss": fde_loss}

    def __str__(self):
        return "Clifford Shared Simplicial MPNN for NBA Dataset"
