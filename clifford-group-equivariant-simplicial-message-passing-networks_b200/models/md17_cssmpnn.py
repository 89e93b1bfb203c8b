"""MD17 atomic-motion model with the reference's interface (csmpn/models/md17_cssmpnn.py:11-178)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..algebra.cliffordalgebra import CliffordAlgebra
from ._shared import Loss, MetricCollection, SharedSimplicialBase, global_mean_pool
from .cegnn_utils import CEMLP, EGCL, MVLinear


class CliffordSharedSimplicialMPNN_md17(SharedSimplicialBase):
    def __init__(self, max_dim: int = 2, num_input: int = 30, num_hidden: int = 32, num_out: int = 10, num_layers: int = 5,
                 condition=True) -> None:
        super().__init__()
        alg = self.algebra = CliffordAlgebra((1, 1, 1))
        self.max_dim, self.condition = max_dim, condition
        self.num_input, self.num_hidden = num_input, num_hidden
        T = self.num_node_type = max_dim + 1 if condition else 0
        self.feature_embedding = MVLinear(alg, num_hidden + T, num_hidden, subspaces=False)
        self.cl_feature_embedding = nn.ModuleList(
            [MVLinear(alg, num_input, num_hidden, subspaces=False)]
            + [CEMLP(alg, (i + 1) * num_input, num_hidden, num_hidden, n_layers=i, normalization_init=0)
               for i in range(1, max_dim + 1)])
        self.sim_type_embedding = nn.Embedding(max_dim + 1, max_dim + 1)
        self.layers = nn.ModuleList([
            EGCL(alg, num_hidden, num_hidden, num_hidden, edge_attr_features=2 * T, node_attr_features=T, aggr="sum",
                 normalization_init=0) for _ in range(num_layers)])
        self.projection = nn.Sequential(CEMLP(alg, num_hidden, num_hidden, num_hidden, n_layers=1),
                                        MVLinear(alg, num_hidden, num_out))
        self.train_metrics, self.valid_metrics, self.test_metrics = (self._setup_metrics() for _ in range(3))
        self.loss_func = nn.MSELoss(reduction="none")

    def _setup_metrics(self):
        return MetricCollection({"loss": Loss(), "ade_loss": Loss(), "fde_loss": Loss()})

    vertex_feature_types = 3

    def vertex_features(self, graph, verts):
        rows, k = verts.shape
        pos = self.grade1(graph.pos[verts].reshape(rows, -1, 3))         # k * frames channels, vertex-major
        vel = self.grade1(graph.vel[verts].reshape(rows, -1, 3))
        chg = self.algebra.embed_grade(graph.charges[verts].reshape(rows, -1, 1), 0)
        return torch.cat((pos, vel, chg), dim=1)

    def featurization(self, x, node_attr):
        return self.feature_embedding(torch.cat((x, node_attr), dim=1))

    def mean_pos(self, loc_node, graph, batch_size, num_frames):
        """centre of every molecule over atoms and frames, broadcast to all of its simplices (md17_cssmpnn.py:140-144)"""
        rows0 = self.simplex_rows(graph)[0]
        per_graph = global_mean_pool(loc_node.reshape(-1, num_frames * 3), graph.batch[rows0], batch_size)
        per_graph = per_graph.reshape(batch_size, num_frames, 3).mean(dim=1, keepdim=True).expand(-1, num_frames, -1)
        return per_graph[graph.x_ind_batch]

    def forward(self, graph, step, mode):
        self.begin_forward(graph)
        batch_size = graph.ptr.shape[0] - 1
        n_real = int(getattr(graph, "n_real_graphs", batch_size))  # shape-padded batches end with one dummy complex
        num_frames = graph.loc.shape[1]
        rows0 = self.simplex_rows(graph)[0]
        loc_node = graph.loc[rows0]
        graph.pos = graph.loc - self.mean_pos(loc_node, graph, batch_size, num_frames)
        node_attr, edge_attr = self.embed_simplex_types(graph)
        x = self.embed_simplicial_complex(graph)
        x = self.featurization(x, node_attr)
        x = self.run_layers(x, graph, edge_attr, node_attr)
        pred = self.projection(x[rows0])[..., 1:4]
        loc_pred = loc_node + pred
        targets = graph.y
        if n_real != batch_size:
            # the dummy complex (last rows) leaves BEFORE the losses: sqrt has no gradient at its all-zero prediction
            v_real = loc_pred.shape[0] // batch_size * n_real
            loc_pred, targets = loc_pred[:v_real], targets[:v_real]
            batch_size = n_real
        sq = F.mse_loss(loc_pred.reshape(-1, 3), targets.reshape(-1, 3), reduction="none")
        ade_loss = torch.sqrt(sq.sum(dim=-1)).reshape(batch_size, -1, num_frames).mean(dim=-1).mean(dim=-1)
        fde_loss = torch.sqrt(F.mse_loss(loc_pred[:, -1, :], targets[:, -1, :], reduction="none").sum(dim=-1)
                              ).reshape(batch_size, -1).mean(dim=-1)
        loss = sq.reshape(batch_size, -1, 3).sum(-1).mean(-1)
        return loss.mean(), {"loss": loss, "ade_loss": ade_loss, "fde_loss": fde_loss}

    def __str__(self):
        return "Clifford Shared Simplicial MPNN for MD17 Dataset"
