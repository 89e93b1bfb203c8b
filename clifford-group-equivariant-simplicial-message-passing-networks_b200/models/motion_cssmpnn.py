"""CMU human-motion model with the reference's interface (csmpn/models/motion_cssmpnn.py:12-170)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..algebra.cliffordalgebra import CliffordAlgebra
from ._shared import Loss, MetricCollection, SharedSimplicialBase
from .cegnn_utils import CEMLP, EGCL, MVLinear


class MotionCliffordSharedSimplicialMPNN(SharedSimplicialBase):
    def __init__(self, max_dim: int = 2, num_input: int = 2, num_hidden: int = 16, num_out: int = 1, num_layers: int = 4,
                 condition=True):
        super().__init__()
        alg = self.algebra = CliffordAlgebra((1, 1, 1))
        self.max_dim, self.condition = max_dim, condition
        T = self.num_node_type = max_dim + 1 if condition else 0
        self.num_input, self.num_hidden = num_input, num_hidden
        self.feature_embedding = MVLinear(alg, num_input + T, num_hidden, subspaces=False)   # unused by forward, as upstream
        self.cl_feature_embedding = nn.ModuleList(
            [MVLinear(alg, num_input, num_hidden, subspaces=False)]
            + [CEMLP(alg, (i + 1) * num_input, num_hidden, num_hidden, n_layers=i, normalization_init=0)
               for i in range(1, max_dim + 1)])
        self.sim_type_embedding = nn.Embedding(max_dim + 1, max_dim + 1)
        self.layers = nn.ModuleList([
            EGCL(alg, num_hidden, num_hidden, num_hidden, edge_attr_features=2 * T, node_attr_features=T, aggr="mean",
                 normalization_init=0) for _ in range(num_layers)])
        self.projection = nn.Sequential(MVLinear(alg, num_hidden, num_out))
        self.train_metrics, self.test_metrics = self._setup_metrics(), self._setup_metrics()
        self.loss_func = nn.MSELoss(reduction="none")

    def _setup_metrics(self):
        return MetricCollection({"loss": Loss()})

    vertex_feature_types = 2

    def vertex_features(self, graph, verts):
        return torch.cat((self.grade1(graph.pos[verts]), self.grade1(graph.vel[verts])), dim=1)

    def forward(self, graph, step, mode):
        batch_size = graph.ptr.shape[0] - 1
        rows0 = self.simplex_rows(graph)[0]
        node_pos = graph.pos[rows0].reshape(batch_size, -1, self.algebra.dim)
        centred = node_pos - node_pos.mean(dim=1, keepdim=True)
        graph.pos = graph.pos.index_copy(0, rows0, centred.reshape(-1, 3))   # the reference centres in place (:146)
        node_attr, edge_attr = self.embed_simplex_types(graph)
        x = self.embed_simplicial_complex(graph)
        x = self.run_layers(x, graph, edge_attr, node_attr)
        pred = self.projection(x[rows0])[..., 0, 1:4]
        pred = node_pos.reshape(-1, self.algebra.dim) + pred
        loss = F.mse_loss(pred, graph.y.reshape(-1, 3), reduction="none").mean(dim=1)
        return loss.mean(), {"loss": loss}

    def __str__(self):
        return "Clifford Shared Simplicial MPNN for Motion Dataset"
