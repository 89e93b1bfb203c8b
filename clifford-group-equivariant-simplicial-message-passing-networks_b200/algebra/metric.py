"""Blade ordering and the sparse Cayley (geometric multiplication) table of Cl(p,q,r).

Mirrors the interface of the reference's ``csmpn/algebra/metric.py`` (``ShortLexBasisBladeOrder``
:18-29, ``canonical_reordering_sign[_euclidean]`` :50-79, ``gmt_element`` :82-89, ``construct_gmt``
:92-120) but is written around one closed form instead of nested Python loops:

* a basis blade is a bitmap over the ``n`` basis vectors; blades are ordered grade-major and, inside
  a grade, by the lexicographic order of their (ascending) basis-vector tuples ("short-lex");
* ``e_A e_B = s(A,B) * prod_{v in A&B} metric[v] * e_{A^B}`` where ``s`` is the parity of the number
  of pairs ``(i in A, j in B)`` with ``i > j`` (each such pair is one transposition).

``product_table`` is the single source of truth for every table the CUDA kernels use
(``csrc/gen_algebra.py`` turns it into unrolled device code).
"""
from __future__ import annotations

import numpy as np
import torch


def _bits(bitmap: int):
    return tuple(i for i in range(bitmap.bit_length()) if (bitmap >> i) & 1)


def blade_bitmaps(n_vectors: int):
    """Bitmaps of all 2**n blades in short-lex order (python ints)."""
    return sorted(range(1 << n_vectors), key=lambda b: (bin(b).count("1"), _bits(b)))


class ShortLexBasisBladeOrder:
    """index <-> bitmap maps and per-blade grades (reference ``metric.py:18-29``)."""

    def __init__(self, n_vectors: int):
        order = blade_bitmaps(n_vectors)
        self.index_to_bitmap = torch.tensor(order, dtype=torch.int64)
        self.grades = torch.tensor([bin(b).count("1") for b in order], dtype=torch.int64)
        inv = [0] * len(order)
        for i, b in enumerate(order):
            inv[b] = i
        self.bitmap_to_index = torch.tensor(inv, dtype=torch.int64)


def count_set_bits(bitmap: int) -> int:
    return bin(int(bitmap)).count("1")


def canonical_reordering_sign_euclidean(bitmap_a: int, bitmap_b: int) -> int:
    """(-1)**(#pairs i in A, j in B with i > j)  (reference ``metric.py:50-62``)."""
    a, b = int(bitmap_a), int(bitmap_b)
    swaps = 0
    for j in _bits(b):
        swaps += count_set_bits(a >> (j + 1))
    return -1 if (swaps & 1) else 1


def canonical_reordering_sign(bitmap_a: int, bitmap_b: int, metric):
    """Reordering sign times the metric of the contracted basis vectors (``metric.py:65-79``)."""
    out = canonical_reordering_sign_euclidean(bitmap_a, bitmap_b)
    for v in _bits(int(bitmap_a) & int(bitmap_b)):
        out = out * metric[v]
    return out


def gmt_element(bitmap_a: int, bitmap_b: int, sig_array):
    """(output bitmap, coefficient) of the product of two basis blades (``metric.py:82-89``)."""
    return int(bitmap_a) ^ int(bitmap_b), canonical_reordering_sign(bitmap_a, bitmap_b, sig_array)


def product_table(metric):
    """Sparse Cayley structure of Cl(metric).

    Returns a dict of numpy arrays with B = 2**dim, G = dim + 1:
      bitmap[B], grade[B], index_of_bitmap[B],
      out[B,B]    -- blade index j of e_i e_k,
      coef[B,B]   -- float64 coefficient c[i,j(i,k),k],
      sign[B,B]   -- the Euclidean reordering sign only (int8),
      common[B,B] -- bitmap_i & bitmap_k (which metric entries were contracted),
      mfac[B]     -- prod of metric over the bits of a *bitmap* (index = bitmap),
      qsign[B]    -- beta_i * c[i,0,i]: per-blade sign of the quadratic form (``cliffordalgebra.py:69-71,119-146``),
      paths[G,G,G] bool, path_index[G,G,G] (row-major rank of the True entries, -1 elsewhere).
    """
    metric = [float(m) for m in metric]
    dim = len(metric)
    B = 1 << dim
    order = blade_bitmaps(dim)
    inv = np.zeros(B, dtype=np.int64)
    for i, b in enumerate(order):
        inv[b] = i
    grade = np.array([bin(b).count("1") for b in order], dtype=np.int64)
    mfac = np.ones(B, dtype=np.float64)
    for bm in range(B):
        for v in _bits(bm):
            mfac[bm] *= metric[v]
    out = np.zeros((B, B), dtype=np.int64)
    sign = np.zeros((B, B), dtype=np.int8)
    common = np.zeros((B, B), dtype=np.int64)
    coef = np.zeros((B, B), dtype=np.float64)
    for i, bi in enumerate(order):
        for k, bk in enumerate(order):
            out[i, k] = inv[bi ^ bk]
            sign[i, k] = canonical_reordering_sign_euclidean(bi, bk)
            common[i, k] = bi & bk
            coef[i, k] = sign[i, k] * mfac[bi & bk]
    beta = np.array([(-1.0) ** (g * (g - 1) // 2) for g in grade])
    qsign = beta * np.array([coef[i, i] for i in range(B)])
    G = dim + 1
    paths = np.zeros((G, G, G), dtype=bool)
    for i in range(B):
        for k in range(B):
            if coef[i, k] != 0:
                paths[grade[i], grade[out[i, k]], grade[k]] = True
    path_index = -np.ones((G, G, G), dtype=np.int64)
    path_index[paths] = np.arange(int(paths.sum()))
    return dict(
        dim=dim, bitmap=np.array(order, dtype=np.int64), grade=grade, index_of_bitmap=inv,
        out=out, coef=coef, sign=sign, common=common, mfac=mfac, qsign=qsign,
        paths=paths, path_index=path_index,
    )


def construct_gmt(index_to_bitmap, bitmap_to_index, signature):
    """Sparse COO Cayley tensor with coords (i_left, j_out, k_right) (``metric.py:92-120``)."""
    n = len(index_to_bitmap)
    sig = [float(s) for s in signature]
    tab = product_table(sig)
    ii, kk = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    coords = torch.tensor(np.stack([ii.ravel(), tab["out"].ravel(), kk.ravel()]), dtype=torch.int64)
    vals = torch.tensor(tab["coef"].ravel(), dtype=torch.float32)
    return torch.sparse_coo_tensor(indices=coords, values=vals, size=(n, n, n), check_invariants=False)
