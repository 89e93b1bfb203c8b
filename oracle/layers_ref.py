"""Plain-PyTorch CPU restatement of the CSMPN hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference lines it restates (paths relative to the reference root).  The code is
written from the formulas with explicit (i, k) -> (j, coefficient) tables; it shares no
code with the product package and none with the reference.  It works in fp32 and fp64 and is differentiable
by autograd, so it serves as the parity oracle for forward values and for every gradient.

Parameters are passed as flat dicts keyed exactly like the reference's ``state_dict`` (e.g.
``"edge_model.layers.0.2.linear_right.weight"``), so the same dict drives the reference, this oracle and the
product modules.
"""
from __future__ import annotations

import itertools
import math

import torch

EPS = 1e-6  # csmpn/models/cegnn_utils.py:5


class RefAlgebra:
    """Tables of Cl(metric): csmpn/algebra/metric.py:18-120, cliffordalgebra.py:11-42,238-252."""

    def __init__(self, metric):
        self.metric = [float(m) for m in metric]
        self.dim = d = len(self.metric)
        self.B = 1 << d
        self.G = d + 1
        # short-lex blade order: by grade, then lexicographic in the basis-vector tuple (metric.py:18-29)
        blades = []
        for g in range(d + 1):
            for combo in itertools.combinations(range(d), g):
                blades.append(sum(1 << v for v in combo))
        self.bitmap = blades
        index_of = {b: i for i, b in enumerate(blades)}
        self.grade = [bin(b).count("1") for b in blades]
        I, J, K, C = [], [], [], []
        for i, a in enumerate(blades):
            for k, b in enumerate(blades):
                # reordering sign: parity of #(p in a, q in b, p > q)  (metric.py:50-62)
                swaps = sum(1 for p in range(d) for q in range(d) if (a >> p) & 1 and (b >> q) & 1 and p > q)
                c = -1.0 if swaps % 2 else 1.0
                for v in range(d):  # contracted vectors bring their metric entry (metric.py:65-79)
                    if (a & b) >> v & 1:
                        c *= self.metric[v]
                I.append(i), J.append(index_of[a ^ b]), K.append(k), C.append(c)
        self.I = torch.tensor(I)
        self.J = torch.tensor(J)
        self.K = torch.tensor(K)
        self.C = torch.tensor(C, dtype=torch.float64)
        self.grade_t = torch.tensor(self.grade)
        self.subspaces = [self.grade.count(g) for g in range(self.G)]
        # beta_i * c[i,0,i]  (cliffordalgebra.py:69-71, 119-146)
        beta = [(-1.0) ** (g * (g - 1) // 2) for g in self.grade]
        self.qsign = torch.tensor([beta[i] * C[i * self.B + i] for i in range(self.B)], dtype=torch.float64)
        # grade paths in row-major (g_i, g_j, g_k) order (cliffordalgebra.py:238-252, cegnn_utils.py:133)
        paths = torch.zeros(self.G, self.G, self.G, dtype=torch.bool)
        for i, j, k, c in zip(I, J, K, C):
            if c != 0:
                paths[self.grade[i], self.grade[j], self.grade[k]] = True
        self.paths = paths
        pid = -torch.ones(self.G, self.G, self.G, dtype=torch.long)
        pid[paths] = torch.arange(int(paths.sum()))
        self.path_of_term = pid[self.grade_t[self.I], self.grade_t[self.J], self.grade_t[self.K]]
        self.n_paths = int(paths.sum())

    def dense_cayley(self, dtype=torch.float32):
        c = torch.zeros(self.B, self.B, self.B, dtype=dtype)
        c[self.I, self.J, self.K] = self.C.to(dtype)
        return c


def geometric_product(alg: RefAlgebra, a, b):
    """out_j = sum_{i,k} a_i c[i,j,k] b_k   (cliffordalgebra.py:44-54)"""
    a, b = torch.broadcast_tensors(a, b)
    terms = a[..., alg.I] * b[..., alg.K] * alg.C.to(a.dtype)
    out = torch.zeros_like(a)
    return out.index_add(-1, alg.J, terms)


def grade_q(alg: RefAlgebra, x):
    """[..., B] -> [..., G]: q_g = sum_{i in g} beta_i c[i,0,i] x_i^2   (cliffordalgebra.py:143-146,162-168)"""
    sq = x * x * alg.qsign.to(x.dtype)
    out = torch.zeros(*x.shape[:-1], alg.G, dtype=x.dtype)
    return out.index_add(-1, alg.grade_t, sq)


def smooth_abs_sqrt(q, eps=1e-16):
    """(q^2 + eps)^(1/4)   (cliffordalgebra.py:148-149)"""
    return (q * q + eps) ** 0.25


def mvlinear(alg, x, weight, bias=None):
    """y[r,n,i] = sum_m x[r,m,i] W[n,m,g(i)] + bias[n] [i == 0]   (cegnn_utils.py:326-338)"""
    if weight.dim() == 3:
        w = weight[:, :, alg.grade_t]  # [n, m, B]
        y = torch.einsum("rmi,nmi->rni", x, w)
    else:
        y = torch.einsum("rmi,nm->rni", x, weight)
    if bias is not None:
        y = y.clone()
        y[..., 0] = y[..., 0] + bias.reshape(1, -1)
    return y


def mvsilu(alg, x, a, b):
    """y_i = sigmoid(a[n,g] inv_g + b[n,g]) x_i, inv_0 = x_0, inv_g = q_g   (cegnn_utils.py:73-83)"""
    inv = grade_q(alg, x)
    inv = torch.cat([x[..., :1], inv[..., 1:]], dim=-1)
    gate = torch.sigmoid(a.reshape(1, -1, alg.G) * inv + b.reshape(1, -1, alg.G))
    return gate[..., alg.grade_t] * x


def normalization(alg, x, a):
    """y_i = x_i / (sigmoid(a[n,g]) (norm_g - 1) + 1 + EPS)   (cegnn_utils.py:42-51)"""
    norms = smooth_abs_sqrt(grade_q(alg, x))
    den = torch.sigmoid(a).reshape(1, -1, alg.G) * (norms - 1) + 1 + EPS
    return x / den[..., alg.grade_t]


def mvlayernorm(alg, x, a):
    """y = a[n] x / (mean_n (Q^2 + 1e-16)^(1/4) + EPS), Q over all blades   (cegnn_utils.py:93-96)"""
    Q = (x * x * alg.qsign.to(x.dtype)).sum(-1, keepdim=True)
    mu = smooth_abs_sqrt(Q).mean(dim=1, keepdim=True) + EPS
    return a.reshape(1, -1, 1) * x / mu


def weighted_gp_tables(alg, x, r, w):
    """z[r,n,j] = sum_{i,k} x_i c[i,j,k] w[n, path(g_i,g_j,g_k)] r_k   (cegnn_utils.py:126-140,151) -- term by term"""
    coef = alg.C.to(x.dtype)
    keep = alg.path_of_term >= 0
    I, J, K = alg.I[keep], alg.J[keep], alg.K[keep]
    wt = w[:, alg.path_of_term[keep]] * coef[keep]  # [C, terms]
    terms = x[..., I] * r[..., K] * wt.unsqueeze(0)
    out = torch.zeros_like(x)
    return out.index_add(-1, J, terms)


def weighted_gp(alg, x, r, w):
    """Same contraction evaluated the way the reference does it on CPU: scatter the path weights into a dense
    [C,B,B,B] tensor and einsum (cegnn_utils.py:126-140,151).  Used by the timed CPU baseline so that the baseline
    runs the reference's algorithm, not a slower table walk; tests check it equals ``weighted_gp_tables``."""
    keep = alg.path_of_term >= 0
    dense = torch.zeros(w.shape[0], alg.B, alg.B, alg.B, dtype=x.dtype)
    dense[:, alg.I[keep], alg.J[keep], alg.K[keep]] = w[:, alg.path_of_term[keep]] * alg.C.to(x.dtype)[keep]
    return torch.einsum("bni,nijk,bnk->bnj", x, dense, r)


def sgp(alg, x, p, prefix):
    """SteerableGeometricProductLayer.forward (cegnn_utils.py:142-155), include_first_order=True"""
    right = mvlinear(alg, x, p[prefix + "linear_right.weight"])
    right = normalization(alg, right, p[prefix + "normalization.a"])
    left = mvlinear(alg, x, p[prefix + "linear_left.weight"], p[prefix + "linear_left.bias"])
    return (left + weighted_gp(alg, x, right, p[prefix + "weight"])) / math.sqrt(2)


def cemlp_block(alg, x, p, prefix):
    """one [MVLinear, MVSiLU, SGP, MVLayerNorm] block of CEMLP (cegnn_utils.py:180-207); prefix = 'layers.k.'"""
    x = mvlinear(alg, x, p[prefix + "0.weight"], p.get(prefix + "0.bias"))
    x = mvsilu(alg, x, p[prefix + "1.a"], p[prefix + "1.b"])
    x = sgp(alg, x, p, prefix + "2.")
    return mvlayernorm(alg, x, p[prefix + "3.a"])


def cemlp(alg, x, p, prefix=""):
    """CEMLP.forward (cegnn_utils.py:210-213)"""
    k = 0
    while (prefix + f"layers.{k}.0.weight") in p:
        x = cemlp_block(alg, x, p, prefix + f"layers.{k}.")
        k += 1
    return x


def egcl(alg, h, edge_index, edge_attr, node_attr, p, aggr="mean", residual=True, prefix=""):
    """EGCL.forward with PyG source_to_target propagate (cegnn_utils.py:254-284).

    message: CEMLP(cat(h_i - h_j, edge_attr)); aggregate at edge_index[1] (sum | mean with count clamped to 1);
    update: h + CEMLP(cat(h, agg, node_attr)).
    """
    src, dst = edge_index[0], edge_index[1]
    inp = h[dst] - h[src]
    if edge_attr is not None:
        inp = torch.cat([inp, edge_attr], dim=1)
    msg = cemlp(alg, inp, p, prefix + "edge_model.")
    agg = torch.zeros_like(h[:, : msg.shape[1]]).index_add(0, dst, msg)
    if aggr == "mean":
        cnt = torch.zeros(h.shape[0], dtype=h.dtype).index_add(0, dst, torch.ones(dst.shape[0], dtype=h.dtype))
        agg = agg / cnt.clamp(min=1).reshape(-1, 1, 1)
    parts = [h, agg] + ([node_attr] if node_attr is not None else [])
    out = cemlp(alg, torch.cat(parts, dim=1), p, prefix + "node_model.")
    return h + out if residual else out


# ------------------------------------------------------------------------------------------------
def init_cemlp_params(alg, c_in, c_hidden, c_out, n_layers=2, gen=None, prefix="", dtype=torch.float32):
    """Random parameters with the reference's names, shapes and initial distributions
    (cegnn_utils.py:40, 60-61, 91, 124, 301-323), perturbed away from the all-ones / all-zeros defaults so that
    every gradient path is exercised."""
    G, P = alg.G, alg.n_paths
    p = {}

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=gen, dtype=torch.float64) * std).to(dtype)

    cin = c_in
    for k in range(n_layers):
        c = c_hidden if k < n_layers - 1 else c_out
        pre = f"{prefix}layers.{k}."
        p[pre + "0.weight"] = rn(c, cin, G, std=1 / math.sqrt(cin))
        p[pre + "0.bias"] = rn(1, c, 1, std=0.1)
        p[pre + "1.a"] = 1 + rn(1, c, G, std=0.1)
        p[pre + "1.b"] = rn(1, c, G, std=0.1)
        p[pre + "2.weight"] = rn(c, P, std=1 / math.sqrt(alg.dim + 1))
        p[pre + "2.normalization.a"] = rn(c, G, std=0.3)
        p[pre + "2.linear_right.weight"] = rn(c, c, G, std=1 / math.sqrt(c))
        p[pre + "2.linear_left.weight"] = rn(c, c, G, std=1 / math.sqrt(c))
        p[pre + "2.linear_left.bias"] = rn(1, c, 1, std=0.1)
        p[pre + "3.a"] = 1 + rn(1, c, std=0.1)
        cin = c
    return p


def init_egcl_params(alg, c, t, gen=None, dtype=torch.float32):
    """Parameters of EGCL(in=hidden=out=c, edge_attr_features=2t, node_attr_features=t)."""
    p = init_cemlp_params(alg, c + 2 * t, c, c, 2, gen, "edge_model.", dtype)
    p.update(init_cemlp_params(alg, 2 * c + t, c, c, 2, gen, "node_model.", dtype))
    return p
