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
    def sandwich(self, u, v, w):
        return self.geometric_product(self.geometric_product(u, v), w)

    def output_blades(self, blades_left, blades_right):
        blades = []
        for blade_left in blades_left:
            for blade_right in blades_right:
                bitmap_left = self.bbo.index_to_bitmap[blade_left]
                bitmap_right = self.bbo.index_to_bitmap[blade_right]
                bitmap_out, _ = gmt_element(int(bitmap_left), int(bitmap_right), self._metric_list)
                blades.append(int(self.bbo.bitmap_to_index[bitmap_out]))
        return torch.tensor(blades)

    def random(self, n=None):
        if n is None:
            n = 1
        return torch.randn(n, self.n_blades)

    def random_vector(self, n=None):
        if n is None:
            n = 1
        vector_indices = self.bbo_grades == 1
        v = torch.zeros(n, self.n_blades, device=self.cayley.device)
        v[:, vector_indices] = torch.randn(n, int(vector_indices.sum()), device=self.cayley.device)
        return v

    def parity(self, mv):
        is_odd = torch.all(mv[..., self.even_grades] == 0)
        is_even = torch.all(mv[..., self.odd_grades] == 0)
        if is_odd ^ is_even:
            return is_odd
        else:
            raise ValueError("This is not a homogeneous element.")

    def eta(self, w):
        return (-1) ** self.parity(w)

    def alpha_w(self, w, mv):
        return self.even_grades * mv + self.eta(w) * self.odd_grades * mv

    def inverse(self, mv, blades=None):
        # kept as in the reference (cliffordalgebra.py:215-217), including its normalisation quirk
        mv_ = self.beta(mv, blades=blades)
        return mv_ / self.b(mv, mv_)

    def rho(self, w, mv):
        return self.sandwich(w, self.alpha_w(w, mv), self.inverse(w))

    def reduce_geometric_product(self, inputs):
        return functools.reduce(self.geometric_product, inputs)

    def versor(self, order=None, normalized=True):
        if order is None:
            order = self.dim if self.dim % 2 == 0 else self.dim - 1
        vectors = self.random_vector(order)
        versor = self.reduce_geometric_product(vectors[:, None])
        if normalized:
            versor = versor / self.norm(versor)[..., :1]
        return versor

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
