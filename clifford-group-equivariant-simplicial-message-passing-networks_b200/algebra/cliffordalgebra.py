"""CliffordAlgebra with the reference's public surface (csmpn/algebra/cliffordalgebra.py:10-262), computing on
B200 through the table-driven kernels in csrc/ (C ABI in include/csmpn_b200.h).

Differences that matter:
  * ``geometric_product`` is a sign/index-table kernel over the B^2 non-zeros of the Cayley tensor
    (cliffordalgebra.py:44-54 contracts the dense [B,B,B] tensor with einsum);
  * ``qs`` / ``norms`` / ``norm`` / ``q`` are signed sums of squares computed in one pass
    (cliffordalgebra.py:119-168 route them through geometric_product with blade subsets);
  * all compute requires CUDA fp32 tensors; pure views (get_grade, embed, split, ...) work anywhere.
Buffers (``metric, subspaces, bbo_grades, even_grades, odd_grades, cayley``) keep the reference's names and
values so a reference ``state_dict`` loads unchanged.
"""
from __future__ import annotations

import functools
import itertools
import math

import torch
from torch import nn

from .. import _lib
from .metric import ShortLexBasisBladeOrder, construct_gmt, gmt_element, product_table


class _GPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, dim, metric, shape):
        a2 = _lib.f32c(a).reshape(-1, a.shape[-1])
        b2 = _lib.f32c(b).reshape(-1, b.shape[-1])
        n = max(a2.shape[0], b2.shape[0])
        a_b = int(a2.shape[0] == 1 and n > 1)
        b_b = int(b2.shape[0] == 1 and n > 1)
        out = torch.empty((n, a2.shape[-1]), dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib().csmpn_gp_fwd(dim, metric, _lib.ptr(a2), _lib.ptr(b2), _lib.ptr(out), n, a_b, b_b,
                                           _lib.stream_ptr(a.device)), "gp_fwd")
        ctx.save_for_backward(a2, b2)
        ctx.meta = (dim, metric, a.shape, b.shape, a_b, b_b)
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, go):
        a2, b2 = ctx.saved_tensors
        dim, metric, ashape, bshape, a_b, b_b = ctx.meta
        go2 = _lib.f32c(go).reshape(-1, go.shape[-1])
        n = go2.shape[0]
        ga = torch.empty_like(go2)
        gb = torch.empty_like(go2)
        _lib.check(_lib.lib().csmpn_gp_bwd(dim, metric, _lib.ptr(a2), _lib.ptr(b2), _lib.ptr(go2), _lib.ptr(ga),
                                           _lib.ptr(gb), n, a_b, b_b, _lib.stream_ptr(go.device)), "gp_bwd")
        ga = ga.sum(0, keepdim=True) if a_b else ga
        gb = gb.sum(0, keepdim=True) if b_b else gb
        return ga.reshape(ashape), gb.reshape(bshape), None, None, None


class _FormsFn(torch.autograd.Function):
    """per-grade q (mode 0) or smooth norm (mode 1): [..., B] -> [..., G]"""

    @staticmethod
    def forward(ctx, x, dim, metric, mode):
        x2 = _lib.f32c(x).reshape(-1, x.shape[-1])
        out = torch.empty((x2.shape[0], dim + 1), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().csmpn_grade_forms_fwd(dim, metric, _lib.ptr(x2), _lib.ptr(out), x2.shape[0], mode,
                                                    _lib.stream_ptr(x.device)), "grade_forms_fwd")
        ctx.save_for_backward(x2)
        ctx.meta = (dim, metric, mode, x.shape)
        return out.reshape(*x.shape[:-1], dim + 1)

    @staticmethod
    def backward(ctx, go):
        (x2,) = ctx.saved_tensors
        dim, metric, mode, shape = ctx.meta
        go2 = _lib.f32c(go).reshape(-1, dim + 1)
        gx = torch.empty_like(x2)
        _lib.check(_lib.lib().csmpn_grade_forms_bwd(dim, metric, _lib.ptr(x2), _lib.ptr(go2), _lib.ptr(gx),
                                                    x2.shape[0], mode, _lib.stream_ptr(go.device)), "grade_forms_bwd")
        return gx.reshape(shape), None, None, None


class CliffordAlgebra(nn.Module):
    def __init__(self, metric):
        super().__init__()
        self.register_buffer("metric", torch.as_tensor(metric))
        self.num_bases = len(metric)
        self.bbo = ShortLexBasisBladeOrder(self.num_bases)
        self.dim = len(self.metric)
        if not 1 <= self.dim <= 5:
            raise ValueError(f"csmpn_b200 supports algebras of dimension 1..5 (got {self.dim})")
        self.n_blades = len(self.bbo.grades)
        self._metric_list = [float(m) for m in self.metric.tolist()]
        self._metric_c = _lib.metric_host(self._metric_list)
        self._table = product_table(self._metric_list)
        self.is_euclidean = all(m == 1.0 for m in self._metric_list)
        B = self.n_blades
        cayley = torch.zeros(B, B, B, dtype=torch.get_default_dtype())
        ii = torch.arange(B).repeat_interleave(B)
        kk = torch.arange(B).repeat(B)
        cayley[ii, torch.as_tensor(self._table["out"]).reshape(-1), kk] = torch.as_tensor(
            self._table["coef"], dtype=torch.get_default_dtype()).reshape(-1)
        self.grades = self.bbo.grades.unique()
        self.register_buffer("subspaces", torch.tensor(tuple(math.comb(self.dim, int(g)) for g in self.grades)))
        self.n_subspaces = len(self.grades)
        self.grade_to_slice = self._grade_to_slice(self.subspaces)
        self.grade_to_index = [torch.tensor(range(*s.indices(s.stop))) for s in self.grade_to_slice]
        self.register_buffer("bbo_grades", self.bbo.grades.to(torch.get_default_dtype()))
        self.register_buffer("even_grades", self.bbo_grades % 2 == 0)
        self.register_buffer("odd_grades", ~self.even_grades)
        self.register_buffer("cayley", cayley)
        # column of SteerableGeometricProductLayer.weight for each Euclidean grade path (-1: path absent
        # for this metric).  Euclidean metrics: identity.
        eu = product_table([1.0] * self.dim)
        cols = []
        for gi, gj, gk in zip(*eu["paths"].nonzero()):
            cols.append(int(self._table["path_index"][gi, gj, gk]))
        self._euclid_path_cols = cols
        self.n_paths_euclid = len(cols)

    # ---------------------------------------------------------------- products
    def geometric_product(self, a, b, blades=None):
        if blades is not None:
            blades_l, blades_o, blades_r = blades
            assert isinstance(blades_l, torch.Tensor)
            assert isinstance(blades_o, torch.Tensor)
            assert isinstance(blades_r, torch.Tensor)
            # the reference slices the Cayley tensor; equivalently embed, multiply, project
            a = self.embed(a, blades_l.to(a.device))
            b = self.embed(b, blades_r.to(b.device))
            return self.geometric_product(a, b)[..., blades_o.to(a.device)]
        _lib.require_cuda(a, b, what="geometric_product")
        if a.shape[-1] != self.n_blades or b.shape[-1] != self.n_blades:
            raise ValueError("geometric_product expects the blade dimension last")
        shape = torch.broadcast_shapes(a.shape, b.shape)
        if a.shape != shape and a.numel() != self.n_blades:
            a = a.expand(shape)
        if b.shape != shape and b.numel() != self.n_blades:
            b = b.expand(shape)
        return _GPFn.apply(a, b, self.dim, self._metric_c, shape)

    def _grade_to_slice(self, subspaces):
        grade_to_slice = list()
        subspaces = torch.as_tensor(subspaces)
        for grade in self.grades:
            index_start = subspaces[:grade].sum()
            index_end = index_start + math.comb(self.dim, int(grade))
            grade_to_slice.append(slice(index_start, index_end))
        return grade_to_slice

    # ---------------------------------------------------------------- involutions (sign flips)
    @functools.cached_property
    def _alpha_signs(self):
        return torch.pow(-1, self.bbo_grades)

    @functools.cached_property
    def _beta_signs(self):
        return torch.pow(-1, self.bbo_grades * (self.bbo_grades - 1) / 2)

    @functools.cached_property
    def _gamma_signs(self):
        return torch.pow(-1, self.bbo_grades * (self.bbo_grades + 1) / 2)

    def _signed(self, signs, mv, blades):
        signs = signs.to(mv.device)
        if blades is not None:
            signs = signs[blades]
        return signs * mv.clone()

    def alpha(self, mv, blades=None):
        return self._signed(self._alpha_signs, mv, blades)

    def beta(self, mv, blades=None):
        return self._signed(self._beta_signs, mv, blades)

    def gamma(self, mv, blades=None):
        return self._signed(self._gamma_signs, mv, blades)

    def zeta(self, mv):
        return mv[..., :1]

    # ---------------------------------------------------------------- embedding / projection (views, zero fill)
    def embed(self, tensor: torch.Tensor, tensor_index: torch.Tensor) -> torch.Tensor:
        mv = torch.zeros(*tensor.shape[:-1], 2**self.dim, device=tensor.device, dtype=tensor.dtype)
        mv[..., tensor_index] = tensor
        return mv

    def embed_grade(self, tensor: torch.Tensor, grade: int) -> torch.Tensor:
        mv = torch.zeros(*tensor.shape[:-1], 2**self.dim, device=tensor.device)
        s = self.grade_to_slice[grade]
        mv[..., s] = tensor
        return mv

    def get(self, mv: torch.Tensor, blade_index) -> torch.Tensor:
        blade_index = tuple(blade_index)
        return mv[..., blade_index]

    def get_grade(self, mv: torch.Tensor, grade: int) -> torch.Tensor:
        s = self.grade_to_slice[grade]
        return mv[..., s]

    # ---------------------------------------------------------------- bilinear / quadratic forms
    def b(self, x, y, blades=None):
        if blades is not None:
            assert len(blades) == 2
            beta_blades = blades[0]
            blades = (blades[0], torch.tensor([0]), blades[1])
        else:
            blades = torch.tensor(range(self.n_blades))
            blades = (blades, torch.tensor([0]), blades)
            beta_blades = None
        return self.geometric_product(self.beta(x, blades=beta_blades), y, blades=blades)

    def q(self, mv, blades=None):
        if blades is not None:
            blades = (blades, blades)
        return self.b(mv, mv, blades=blades)

    def _smooth_abs_sqrt(self, input, eps=1e-16):
        return (input**2 + eps) ** 0.25

    def norm(self, mv, blades=None):
        return self._smooth_abs_sqrt(self.q(mv, blades=blades))

    def _forms(self, mv, grades, mode):
        """All per-grade forms in one kernel pass; returns the list the reference returns."""
        _lib.require_cuda(mv, what="qs/norms")
        allg = _FormsFn.apply(mv, self.dim, self._metric_c, mode)
        if grades is None:
            grades = self.grades
        return [allg[..., int(g) : int(g) + 1] for g in grades]

    def norms(self, mv, grades=None):
        return self._forms(mv, grades, 1)

    def qs(self, mv, grades=None):
        return self._forms(mv, grades, 0)

    # ---------------------------------------------------------------- versor utilities
    # Off the hot path (no model calls them); kept so that code written against the reference's CliffordAlgebra -- its
    # equivariance checks use rho / versor (cliffordalgebra.py:170-236) -- keeps working.  Same results, own wording.
    def sandwich(self, u, v, w):
        return self.geometric_product(self.geometric_product(u, v), w)

    def output_blades(self, blades_left, blades_right):
        """blade index of e_l e_r for every (l, r), l-major"""
        to_bitmap, to_index = self.bbo.index_to_bitmap, self.bbo.bitmap_to_index
        return torch.tensor([int(to_index[gmt_element(int(to_bitmap[l]), int(to_bitmap[r]), self._metric_list)[0]])
                             for l, r in itertools.product(blades_left, blades_right)])

    def random(self, n=None):
        return torch.randn(1 if n is None else n, self.n_blades)

    def random_vector(self, n=None):
        rows, dev = (1 if n is None else n), self.cayley.device
        grade1 = torch.nonzero(self.bbo_grades == 1).squeeze(1).to(dev)
        draws = torch.randn(rows, grade1.numel(), device=dev)
        return torch.zeros(rows, self.n_blades, device=dev).index_copy_(1, grade1, draws)

    def parity(self, mv):
        """0-dim bool tensor: True for an odd element, False for an even one; anything mixed (or zero) is rejected"""
        no_even_part = bool((mv[..., self.even_grades] == 0).all())
        no_odd_part = bool((mv[..., self.odd_grades] == 0).all())
        if no_even_part == no_odd_part:
            raise ValueError("This is not a homogeneous element.")
        return torch.tensor(no_even_part)

    def eta(self, w):
        return -1 if bool(self.parity(w)) else 1

    def alpha_w(self, w, mv):
        # the grade involution of mv when w is odd, mv itself when w is even
        return torch.where(self.odd_grades.to(mv.device), self.eta(w) * mv, mv)

    def inverse(self, mv, blades=None):
        reverse = self.beta(mv, blades=blades)
        return reverse / self.b(mv, reverse)  # the reference's normalisation (cliffordalgebra.py:215-217), quirk included

    def rho(self, w, mv):
        return self.sandwich(w, self.alpha_w(w, mv), self.inverse(w))

    def reduce_geometric_product(self, inputs):
        factors = iter(inputs)
        acc = next(factors)
        for f in factors:
            acc = self.geometric_product(acc, f)
        return acc

    def versor(self, order=None, normalized=True):
        if order is None:
            order = self.dim - self.dim % 2  # the largest even number of reflections
        v = self.reduce_geometric_product(self.random_vector(order)[:, None])
        return v / self.norm(v)[..., :1] if normalized else v

    def rotor(self):
        return self.versor()

    @functools.cached_property
    def geometric_product_paths(self):
        return torch.as_tensor(self._table["paths"].copy())

    # ---------------------------------------------------------------- reshapes
    def split(self, mv):
        return mv.reshape(mv.shape[0], -1, 2**self.dim)

    def flatten(self, mv):
        return mv.reshape(mv.shape[0], -1)
