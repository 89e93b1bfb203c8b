"""Convex-hull volume model with the reference's interface (csmpn/models/hulls_cssmpnn.py:10-164); Cl(5,0)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..algebra.cliffordalgebra import CliffordAlgebra
from ._shared import Loss, MetricCollection, SharedSimplicialBase, global_mean_pool
from .cegnn_utils import CEMLP, EGCL, MVLinear


class HullsCliffordSharedSimplicialMPNN(SharedSimplicialBase):
    learned_type_embedding = False   # one-hot simplex types (hulls_cssmpnn.py:127-131)

    def __init__(self, in_features=1, hidden_features=28, out_features=1, edge_features_in=0, num_layers=3,
                 normalization_init=0, residual=True, aggr="mean", condition=True, max_dim: int = 2):
        super().__init__()
        self.max_dim = max_dim
        alg = self.algebra = CliffordAlgebra((1.0, 1.0, 1.0, 1.0, 1.0))
        self.hidden_features = self.num_hidden = hidden_features
        self.in_features = in_features
        self.n_layers = num_layers
        T = self.num_node_type = max_dim + 1 if condition else 0
        self.cl_feature_embedding = nn.ModuleList(
            [MVLinear(alg, in_features, hidden_features, subspaces=False)]
            + [CEMLP(alg, (i + 1) * in_features, hidden_features, hidden_features, n_layers=i, normalization_init=0)
               for i in range(1, max_dim + 1)])
        layers = [EGCL(alg, hidden_features, hidden_features, hidden_features, edge_attr_features=2 * T,
                       node_attr_features=T, residual=residual, normalization_init=normalization_init, aggr=aggr)
                  for _ in range(num_layers)]
        self.projection = nn.Sequential(MVLinear(alg, hidden_features, out_features))
        self.readout = nn.Linear(3, 1)   # declared (and unused) upstream; kept so checkpoints load
        self.layers = nn.Sequential(*layers)
        self.train_metrics, self.val_metrics, self.test_metrics = (self._setup_metrics() for _ in range(3))
        self.loss_func = nn.MSELoss(reduction="none")

    def _setup_metrics(self):
        return MetricCollection({"loss": Loss()})

    def _forward(self, h, edges, node_attr=None, edge_attr=None):
        for layer in self.layers:
            h = layer(h, edges, node_attr=node_attr, edge_attr=edge_attr)
        return self.projection(h)

    vertex_feature_types = 1

    def vertex_features(self, graph, verts):
        return self.grade1(graph.input[verts])

    def forward(self, batch, step, mode):
        batch_size = batch.ptr.shape[0] - 1
        rows0 = self.simplex_rows(batch)[0]
        node_pos = batch.input[rows0].reshape(batch_size, -1, self.algebra.dim)
        centred = node_pos - node_pos.mean(dim=1, keepdim=True)
        batch.input = batch.input.index_copy(0, rows0, centred.reshape(-1, self.algebra.dim))
        x = self.embed_simplicial_complex(batch)
        node_attr, edge_attr = self.embed_simplex_types(batch)
        pred = self._forward(x, batch.edge_index, node_attr, edge_attr)[:, :, 0]
        pred = global_mean_pool(pred, batch.x_ind_batch, batch_size)
        loss = F.mse_loss(pred.squeeze(-1), batch.target, reduction="none")
        return loss.mean(0), {"loss": loss}

    def __str__(self):
        return "Clifford Shared Simplicial MPNN for Convex Hulls Dataset"
