"""CUDA-graph replay of the layer step for a fixed complex structure.

One EGCL layer forward + backward is ~35 kernels of 10-200 us each; launched eagerly through autograd the host side
(Python autograd nodes, ctypes calls, allocator) leaves gaps between them that add up to 15-20 % of the step.  A batch
of complexes has a static structure (same CSR, same shapes) for every layer of a model and -- for fixed-topology data
such as the motion skeleton -- for every step, so the launch sequence is captured once and replayed.

``GraphedEGCL`` wraps ``torch.cuda.make_graphed_callables``: forward and backward are captured as two graphs behind
one autograd node, parameters stay live (their gradients are produced by the captured backward), inputs are copied
into the graph's static buffers on every call.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .models.ops import CSRGraph, get_csr


class _BoundLayer(nn.Module):
    """EGCL with its (non-tensor) CSR structure bound, so that the call takes tensors only."""

    def __init__(self, layer: nn.Module, graph: CSRGraph):
        super().__init__()
        self.layer = layer
        self._graph = graph

    def forward(self, h, edge_attr, node_attr):
        return self.layer(h, self._graph, edge_attr, node_attr)


class _BoundPairedLayer(_BoundLayer):
    """the same with edge_attr = PairedNodeAttr(node_attr): (node_attr[src] | node_attr[dst]) gathered by the kernel"""

    def forward(self, h, node_attr):
        from .models.cegnn_utils import PairedNodeAttr

        return self.layer(h, self._graph, PairedNodeAttr(node_attr), node_attr)


class GraphedEGCL:
    """layer(h, graph, edge_attr, node_attr) for a FIXED ``graph`` (CSRGraph or edge_index) and fixed shapes, replayed
    from CUDA graphs.  ``h`` may require grad; ``edge_attr`` / ``node_attr`` follow the sample tensors' requires_grad."""

    def __init__(self, layer: nn.Module, graph, h: torch.Tensor, edge_attr, node_attr: torch.Tensor):
        from .models.cegnn_utils import PairedNodeAttr

        if not h.is_cuda:
            raise ValueError("GraphedEGCL needs CUDA tensors")
        self.paired = isinstance(edge_attr, PairedNodeAttr)
        csr = get_csr(graph, h.shape[0])
        if not isinstance(graph, CSRGraph):  # own the structure: set_graph() rewrites it in place
            csr = CSRGraph(graph.clone(), h.shape[0])
        # build every lazily-created helper structure (sorted views of the CSR) BEFORE capture
        with torch.no_grad():
            layer(h, csr, edge_attr, node_attr)
        torch.cuda.synchronize(h.device)
        if self.paired:
            self.bound = _BoundPairedLayer(layer, csr)
            sample = (h.detach().clone().requires_grad_(True),
                      node_attr.detach().clone().requires_grad_(node_attr.requires_grad))
        else:
            self.bound = _BoundLayer(layer, csr)
            sample = (h.detach().clone().requires_grad_(True),
                      edge_attr.detach().clone().requires_grad_(edge_attr.requires_grad),
                      node_attr.detach().clone().requires_grad_(node_attr.requires_grad))
        # capture runs on a side stream while the parameters' AccumulateGrad nodes may date from the default stream; the
        # mismatch is intended here and torch's warning would repeat on every backward
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        self.fn = torch.cuda.make_graphed_callables(self.bound, sample)

    def set_graph(self, edge_index: torch.Tensor):
        """New complexes with the SAME counts (N simplices, E pairs): the CSR and its sorted views are rebuilt in place
        (a handful of eager launches), the captured graphs read them at replay."""
        self.bound._graph.rebuild_(edge_index)

    def __call__(self, h, edge_attr, node_attr):
        if self.paired:  # edge_attr (a PairedNodeAttr or None) is implied by node_attr
            return self.fn(h, node_attr)
        return self.fn(h, edge_attr, node_attr)


class GraphedLayerStep:
    """Forward + backward + gradient pack of one EGCL layer as ONE CUDA graph over STATIC tensors.

    ``h``, ``node_attr`` and ``cot`` are the tensors the graph reads: a host feeder copies every new batch straight into them
    (no per-call copy into graph-private buffers as with ``torch.cuda.make_graphed_callables``) and a step is a single
    ``replay()`` -- a dozen host calls per step instead of ~40, which is what eight feeder processes on one host need.
    ``set_graph(edge_index)`` rebuilds the CSR of new pairs in place (same counts).  Outputs (static as well): ``y``,
    ``grad_h`` and ``flat`` (every parameter gradient, packed in ``layer.parameters()`` order; ``pack=False``: left in
    ``param_grads``, one tensor per parameter)."""

    def __init__(self, layer: nn.Module, edge_index: torch.Tensor, h: torch.Tensor, node_attr: torch.Tensor, cot: torch.Tensor,
                 warmup: int = 3, pack: bool = True):
        from .models.cegnn_utils import PairedNodeAttr

        if not h.is_cuda:
            raise ValueError("GraphedLayerStep needs CUDA tensors")
        self.h, self.node_attr, self.cot = h, node_attr, cot
        self.csr = CSRGraph(edge_index.clone(), h.shape[0])
        self.params = list(layer.parameters())
        self.flat = torch.empty(sum(p.numel() for p in self.params), dtype=torch.float32, device=h.device)
        views = list(self.flat.split([p.numel() for p in self.params]))

        def run():
            hh = self.h.detach().requires_grad_()
            y = layer(hh, self.csr, PairedNodeAttr(self.node_attr), self.node_attr)
            grads = torch.autograd.grad(y, [hh] + self.params, self.cot)
            if pack:
                torch._foreach_copy_(views, [g.reshape(-1) for g in grads[1:]])
            else:
                self.param_grads = list(grads[1:])  # static tensors as well (the graph's pool)
            return y.detach(), grads[0]

        cur = torch.cuda.current_stream(h.device)
        side = torch.cuda.Stream(h.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):   # builds the lazily created helper structures (sorted views of the CSR, workspaces)
                run()
        cur.wait_stream(side)
        torch.cuda.synchronize(h.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.y, self.grad_h = run()

    def set_graph(self, edge_index: torch.Tensor):
        self.csr.rebuild_(edge_index)

    def __call__(self):
        self.graph.replay()
        return self.y, self.grad_h, self.flat
