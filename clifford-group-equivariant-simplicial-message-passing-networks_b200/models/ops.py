"""torch.autograd bridges from the layer classes to the C ABI (include/csmpn_b200.h).

Every function here validates shape / dtype / device, allocates outputs and workspaces with torch (the library
allocates nothing) and calls one or two C entry points on the current CUDA stream.
"""
from __future__ import annotations

import math
import os

import torch

from .. import _lib
from .._lib import check, f32c, lib, ptr, require_cuda, stream_ptr, workspace


def _rows_channels(x, what):
    if x.dim() != 3:
        raise ValueError(f"{what}: expected [rows, channels, blades], got {tuple(x.shape)}")
    return x.shape[0], x.shape[1]


# --------------------------------------------------------------------------------------- MVLinear
class MVLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, dim, subspaces):
        require_cuda(x, weight, bias, what="MVLinear")
        x = f32c(x)
        weight = f32c(weight)
        rows, c_in = _rows_channels(x, "MVLinear")
        c_out = weight.shape[0]
        y = torch.empty((rows, c_out, x.shape[2]), dtype=torch.float32, device=x.device)
        b = None if bias is None else f32c(bias).reshape(-1)
        check(lib().csmpn_mvlinear_fwd(dim, ptr(x), ptr(weight), ptr(b), ptr(y), rows, c_in, c_out, int(subspaces),
                                       stream_ptr(x.device)), "mvlinear_fwd")
        ctx.save_for_backward(x, weight)
        ctx.meta = (dim, subspaces, bias is not None, None if bias is None else bias.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        dim, subspaces, has_bias, bias_shape = ctx.meta
        gy = f32c(gy)
        rows, c_in = x.shape[0], x.shape[1]
        c_out = weight.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            check(lib().csmpn_mvlinear_bwd_input(dim, ptr(gy), ptr(weight), ptr(gx), rows, c_in, c_out, int(subspaces),
                                                 stream_ptr(x.device)), "mvlinear_bwd_input")
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(weight)
            gbf = torch.empty(c_out, dtype=torch.float32, device=x.device) if has_bias else None
            nbytes = lib().csmpn_mvlinear_bwd_weight_workspace(dim, rows, c_in, c_out)
            ws = workspace(nbytes, x.device)
            check(lib().csmpn_mvlinear_bwd_weight(dim, ptr(x), ptr(gy), ptr(gw), ptr(gbf), rows, c_in, c_out,
                                                  int(subspaces), ptr(ws), ws.numel(), stream_ptr(x.device)),
                  "mvlinear_bwd_weight")
            if has_bias:
                gb = gbf.reshape(bias_shape)
        return gx, gw, gb, None, None


# --------------------------------------------------------------------------------------- row-local layers
class MVSiLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, b, dim, metric):
        require_cuda(x, a, b, what="MVSiLU")
        x = f32c(x)
        rows, ch = _rows_channels(x, "MVSiLU")
        a2, b2 = f32c(a).reshape(ch, dim + 1), f32c(b).reshape(ch, dim + 1)
        y = torch.empty_like(x)
        check(lib().csmpn_mvsilu_fwd(dim, metric, ptr(x), ptr(a2), ptr(b2), ptr(y), rows, ch, stream_ptr(x.device)),
              "mvsilu_fwd")
        ctx.save_for_backward(x, a2, b2)
        ctx.meta = (dim, metric, a.shape, b.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, a2, b2 = ctx.saved_tensors
        dim, metric, ashape, bshape = ctx.meta
        gy = f32c(gy)
        rows, ch = x.shape[0], x.shape[1]
        gx = torch.empty_like(x)
        ga = torch.empty_like(a2)
        gb = torch.empty_like(b2)
        ws = workspace(lib().csmpn_param_grad_workspace(ch * (dim + 1) * 2), x.device)
        check(lib().csmpn_mvsilu_bwd(dim, metric, ptr(x), ptr(a2), ptr(b2), ptr(gy), ptr(gx), ptr(ga), ptr(gb), rows, ch,
                                     ptr(ws), ws.numel(), stream_ptr(x.device)), "mvsilu_bwd")
        return gx, ga.reshape(ashape), gb.reshape(bshape), None, None


class MVNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, dim, metric):
        require_cuda(x, a, what="NormalizationLayer")
        x = f32c(x)
        rows, ch = _rows_channels(x, "NormalizationLayer")
        a2 = f32c(a).reshape(ch, dim + 1)
        y = torch.empty_like(x)
        check(lib().csmpn_mvnorm_fwd(dim, metric, ptr(x), ptr(a2), ptr(y), rows, ch, stream_ptr(x.device)), "mvnorm_fwd")
        ctx.save_for_backward(x, a2)
        ctx.meta = (dim, metric, a.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, a2 = ctx.saved_tensors
        dim, metric, ashape = ctx.meta
        gy = f32c(gy)
        rows, ch = x.shape[0], x.shape[1]
        gx = torch.empty_like(x)
        ga = torch.empty_like(a2)
        ws = workspace(lib().csmpn_param_grad_workspace(ch * (dim + 1)), x.device)
        check(lib().csmpn_mvnorm_bwd(dim, metric, ptr(x), ptr(a2), ptr(gy), ptr(gx), ptr(ga), rows, ch, ptr(ws),
                                     ws.numel(), stream_ptr(x.device)), "mvnorm_bwd")
        return gx, ga.reshape(ashape), None, None


class MVLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, dim, metric):
        require_cuda(x, a, what="MVLayerNorm")
        x = f32c(x)
        rows, ch = _rows_channels(x, "MVLayerNorm")
        a2 = f32c(a).reshape(ch)
        y = torch.empty_like(x)
        check(lib().csmpn_mvlayernorm_fwd(dim, metric, ptr(x), ptr(a2), ptr(y), rows, ch, stream_ptr(x.device)),
              "mvlayernorm_fwd")
        ctx.save_for_backward(x, a2)
        ctx.meta = (dim, metric, a.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, a2 = ctx.saved_tensors
        dim, metric, ashape = ctx.meta
        gy = f32c(gy)
        rows, ch = x.shape[0], x.shape[1]
        gx = torch.empty_like(x)
        ga = torch.empty_like(a2)
        ws = workspace(lib().csmpn_param_grad_workspace(ch), x.device)
        check(lib().csmpn_mvlayernorm_bwd(dim, metric, ptr(x), ptr(a2), ptr(gy), ptr(gx), ptr(ga), rows, ch, ptr(ws),
                                          ws.numel(), stream_ptr(x.device)), "mvlayernorm_bwd")
        return gx, ga.reshape(ashape), None, None


class WeightedGPFn(torch.autograd.Function):
    """out = (left + wgp(x, r; w)) * scale   (left may be None)"""

    @staticmethod
    def forward(ctx, x, r, w, left, scale, dim, metric):
        require_cuda(x, r, w, left, what="SteerableGeometricProductLayer")
        x, r, w = f32c(x), f32c(r), f32c(w)
        rows, ch = _rows_channels(x, "SteerableGeometricProductLayer")
        left_c = None if left is None else f32c(left)
        out = torch.empty_like(x)
        check(lib().csmpn_wgp_fwd(dim, metric, ptr(x), ptr(r), ptr(w), ptr(left_c), float(scale), ptr(out), rows, ch,
                                  stream_ptr(x.device)), "wgp_fwd")
        ctx.save_for_backward(x, r, w)
        ctx.meta = (dim, metric, float(scale), left is not None)
        return out

    @staticmethod
    def backward(ctx, go):
        x, r, w = ctx.saved_tensors
        dim, metric, scale, has_left = ctx.meta
        go = f32c(go)
        rows, ch = x.shape[0], x.shape[1]
        gx = torch.empty_like(x)
        gr = torch.empty_like(r)
        gw = torch.empty_like(w)
        ws = workspace(lib().csmpn_param_grad_workspace(w.numel()), x.device)
        check(lib().csmpn_wgp_bwd(dim, metric, ptr(x), ptr(r), ptr(w), ptr(go), scale, ptr(gx), ptr(gr), ptr(gw), rows,
                                  ch, ptr(ws), ws.numel(), stream_ptr(x.device)), "wgp_bwd")
        gleft = go * scale if has_left else None
        return gx, gr, gw, gleft, None, None, None


# --------------------------------------------------------------------------------------- graph plumbing
class CSRGraph:
    """Receiver- and sender-sorted views of an ``edge_index`` ([2,E] int64; row 0 = sender j, row 1 = receiver i).

    Built once per batch and shared by every EGCL layer of a model (the reference re-derives gather/scatter
    indices inside PyG on every ``propagate`` call).
    """

    def __init__(self, edge_index: torch.Tensor, n_nodes: int):
        require_cuda(edge_index, what="CSRGraph")
        if edge_index.dim() != 2 or edge_index.shape[0] != 2 or edge_index.dtype != torch.int64:
            raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
        self.n_nodes = int(n_nodes)
        self.n_pairs = int(edge_index.shape[1])
        if os.environ.get("CSMPN_CHECK_INDICES") == "1" and self.n_pairs:
            # PyG raises IndexError for simplex ids outside [0, n_nodes); the kernels trust them (one host sync to check)
            lo, hi = int(edge_index.min()), int(edge_index.max())
            if lo < 0 or hi >= int(n_nodes):
                raise IndexError(f"edge_index holds simplex ids in [{lo}, {hi}] but the graph has {int(n_nodes)} simplices")
        ei = edge_index.contiguous()
        self.edge_index = ei
        self.src = ei[0]
        self.dst = ei[1]
        dev = ei.device
        E, N = self.n_pairs, self.n_nodes
        self.rowptr_dst = torch.empty(N + 1, dtype=torch.int32, device=dev)
        self.perm_dst = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self.rowptr_src = torch.empty(N + 1, dtype=torch.int32, device=dev)
        self.perm_src = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self._rebuild_ws = None
        self._build_all(None)
        self._key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, N)

    def rebuild_(self, edge_index: torch.Tensor):
        """Rebuild IN PLACE for a new ``edge_index`` of the same shape: every tensor of this object keeps its address, so
        CUDA graphs captured on it (csmpn_b200.graphs) stay valid.  Same node count; values are copied."""
        require_cuda(edge_index, what="CSRGraph.rebuild_")
        if tuple(edge_index.shape) != tuple(self.edge_index.shape) or edge_index.dtype != torch.int64:
            raise ValueError("CSRGraph.rebuild_: edge_index must keep its shape [2, E] and dtype int64")
        self.edge_index.copy_(edge_index)
        # The dozen small launches of the rebuild (two counting sorts, sorted views, inverse permutation) are captured
        # once and replayed as ONE graph launch: a host-fed loop calls this every step, and with eight ranks per box the
        # per-launch driver cost, not the kernels, is what the step pays for (CSMPN_CSR_GRAPH=0: eager launches).
        sg = getattr(self, "_sorted", None)
        g = getattr(self, "_rebuild_graph", None)
        if g is not None and self._rebuild_sorted is sg:
            g.replay()
            return self
        self._build_all(sg)
        if os.environ.get("CSMPN_CSR_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize(self.edge_index.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._build_all(sg)
            self._rebuild_graph, self._rebuild_sorted = graph, sg
        return self

    def _build_all(self, sg):
        """both CSRs (+ the sorted views of ``sg``): six launches for batches of up to 65 536 simplices (csmpn_csr_build_pair,
        CSMPN_CSR_SMALL=0 disables it), else the sixteen of the general path."""
        dev = self.edge_index.device
        E, N = self.n_pairs, self.n_nodes
        s = stream_ptr(dev)
        if getattr(self, "_rebuild_ws", None) is None:
            self._rebuild_ws = workspace(2 * lib().csmpn_csr_workspace(E, N), dev)
        ws = self._rebuild_ws
        if os.environ.get("CSMPN_CSR_SMALL", "1") != "0":
            st = lib().csmpn_csr_build_pair(ptr(self.src), ptr(self.dst), E, N, ptr(self.rowptr_dst), ptr(self.perm_dst),
                                            ptr(self.rowptr_src), ptr(self.perm_src),
                                            None if sg is None else ptr(sg.src_sorted), None if sg is None else ptr(sg.dst_sorted),
                                            None if sg is None else ptr(sg.rank), ptr(ws), ws.numel(), s)
            if st == 0:
                return
            if st != -3:  # anything but "too many simplices for the one-CTA scan"
                check(st, "csr_build_pair")
        check(lib().csmpn_csr_build(ptr(self.dst), E, N, ptr(self.rowptr_dst), ptr(self.perm_dst), ptr(ws), ws.numel(), s),
              "csr_build(dst)")
        check(lib().csmpn_csr_build(ptr(self.src), E, N, ptr(self.rowptr_src), ptr(self.perm_src), ptr(ws), ws.numel(), s),
              "csr_build(src)")
        if sg is not None:
            sg.refresh_()


_CSR_CACHE: dict = {}
_CSR_CACHE_SIZE = 8


def get_csr(edge_index, n_nodes) -> CSRGraph:
    """Small identity-keyed cache: the same edge_index tensor is reused by all layers of a forward pass."""
    if isinstance(edge_index, CSRGraph):
        return edge_index
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(n_nodes))
    g = _CSR_CACHE.get(key)
    if g is None:
        while len(_CSR_CACHE) >= _CSR_CACHE_SIZE:  # least recently used first; holders of a CSRGraph keep theirs alive
            _CSR_CACHE.pop(next(iter(_CSR_CACHE)))
        g = CSRGraph(edge_index, n_nodes)
    else:
        _CSR_CACHE.pop(key)
    _CSR_CACHE[key] = g  # most recently used last
    return g


class GatherDiffFn(torch.autograd.Function):
    """[N, W] -> [E, W]: h[dst] - h[src]; backward is the deterministic two-CSR scatter."""

    @staticmethod
    def forward(ctx, h, graph: CSRGraph):
        h = f32c(h)
        out = torch.empty((graph.n_pairs, h.shape[1]), dtype=torch.float32, device=h.device)
        check(lib().csmpn_gather_diff(ptr(h), ptr(graph.src), ptr(graph.dst), ptr(out), graph.n_pairs, h.shape[1],
                                      stream_ptr(h.device)), "gather_diff")
        ctx.graph = graph
        ctx.n = h.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        g = f32c(g)
        graph = ctx.graph
        gh = torch.empty((ctx.n, g.shape[1]), dtype=torch.float32, device=g.device)
        check(lib().csmpn_scatter_diff(ptr(g), ptr(graph.rowptr_dst), ptr(graph.perm_dst), ptr(graph.rowptr_src),
                                       ptr(graph.perm_src), ptr(gh), ctx.n, g.shape[1], 0, stream_ptr(g.device)),
              "scatter_diff")
        return gh, None


class SegmentReduceFn(torch.autograd.Function):
    """[E, W] -> [N, W]: sum | mean of the messages of each receiver, fixed order."""

    @staticmethod
    def forward(ctx, msg, graph: CSRGraph, mean: bool):
        msg = f32c(msg)
        out = torch.empty((graph.n_nodes, msg.shape[1]), dtype=torch.float32, device=msg.device)
        check(lib().csmpn_segment_reduce(ptr(msg), ptr(graph.rowptr_dst), ptr(graph.perm_dst), ptr(out), graph.n_nodes,
                                         msg.shape[1], int(mean), stream_ptr(msg.device)), "segment_reduce")
        ctx.graph = graph
        ctx.mean = mean
        return out

    @staticmethod
    def backward(ctx, go):
        go = f32c(go)
        graph = ctx.graph
        gm = torch.empty((graph.n_pairs, go.shape[1]), dtype=torch.float32, device=go.device)
        check(lib().csmpn_segment_expand(ptr(go), ptr(graph.dst), ptr(graph.rowptr_dst), ptr(gm), graph.n_pairs,
                                         go.shape[1], int(ctx.mean), stream_ptr(go.device)), "segment_expand")
        return gm, None, None


INV_SQRT2 = 1.0 / math.sqrt(2.0)
