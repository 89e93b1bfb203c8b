#!/usr/bin/env python
"""Generate ``algebra_gen.cuh``: fully unrolled sign/index-table device code for Euclidean Cl(n,0).

The single source of truth is ``algebra/metric.py::product_table`` (checked against the reference's
``csmpn/algebra/metric.py`` in tests/test_algebra_tables.py).  For every supported dimension this
emits ``template<> struct Alg<DIM>`` with

  B, G, P                       blade / grade / grade-path counts
  grade_of[B], grade_start[G+1] blade -> grade, first blade of each grade (blades are grade-major)
  gp(a,b,o)                     o_j  = sum_{i,k} s(i,k) a_i b_k               (reference cliffordalgebra.py:44-54)
  gp_bwd(a,b,go,ga,gb)          adjoints of gp
  wgp(x,r,w,z)                  z_j += sum_{i,k} s(i,k) w[path(g_i,g_j,g_k)] x_i r_k   (cegnn_utils.py:126-155)
  wgp_bwd(x,r,w,dz,dx,dr,dw)    adjoints of wgp (dx, dr, dw are accumulated into)

Everything is straight-line code over register arrays: the Cayley tensor's B^2 non-zeros out of B^3
become B^2 FFMAs with the sign folded into the instruction's negate modifier.

Run:  python gen_algebra.py   (writes algebra_gen.cuh next to this file; csrc/build.py runs it at build time, the output is not tracked)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import importlib.util

_spec = importlib.util.spec_from_file_location("_metric", os.path.join(os.path.dirname(HERE), "algebra", "metric.py"))
_metric = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_metric)
product_table = _metric.product_table

DIMS = (1, 2, 3, 4, 5)


def sgn(s):
    return "+" if s > 0 else "-"


def emit_dim(dim):
    t = product_table([1.0] * dim)
    B = 1 << dim
    G = dim + 1
    grade = t["grade"]
    out = t["out"]
    sign = t["sign"]
    pidx = t["path_index"]
    common = t["common"]

    def mfx(name, i, k):
        """operand `name[i]`, times the metric factor of the contracted basis vectors when MET."""
        c = int(common[i, k])
        return f"{name}[{i}]" if c == 0 else f"(MET ? {name}[{i}] * mf[{c}] : {name}[{i}])"
    P = int(t["paths"].sum())
    gstart = [0]
    for g in range(G):
        gstart.append(gstart[-1] + int((grade == g).sum()))
    L = []
    A = L.append
    A(f"template <> struct Alg<{dim}> {{")
    A(f"  static constexpr int DIM = {dim}, B = {B}, G = {G}, P = {P};")
    go = " : ".join(f"i < {gstart[g + 1]} ? {g}" for g in range(G - 1)) + f" : {G - 1}"
    gs = " : ".join(f"g == {g} ? {gstart[g]}" for g in range(G)) + f" : {gstart[G]}"
    A(f"  __host__ __device__ static constexpr int grade_of(int i) {{ return {go}; }}")
    A(f"  __host__ __device__ static constexpr int grade_start(int g) {{ return {gs}; }}")
    A("  __host__ __device__ static constexpr int grade_size(int g) { return grade_start(g + 1) - grade_start(g); }")
    # ---- plain geometric product
    A("  template <bool MET> __device__ __forceinline__ static void gp(const float* a, const float* b, const float* mf, float* o) {")
    for j in range(B):
        terms = [(i, k) for i in range(B) for k in range(B) if out[i, k] == j]
        first = True
        for (i, k) in terms:
            if first:
                A(f"    o[{j}] = {'' if sign[i, k] > 0 else '-'}{mfx('a', i, k)} * b[{k}];")
                first = False
            else:
                A(f"    o[{j}] = fmaf({'' if sign[i, k] > 0 else '-'}{mfx('a', i, k)}, b[{k}], o[{j}]);")
    A("  }")
    A("  template <bool MET> __device__ __forceinline__ static void gp_bwd(const float* a, const float* b, const float* go, const float* mf, float* ga, float* gb) {")
    for i in range(B):
        first = True
        for k in range(B):
            j = out[i, k]
            n = "" if sign[i, k] > 0 else "-"
            if first:
                A(f"    ga[{i}] = {n}go[{j}] * {mfx('b', k, i)};")
                first = False
            else:
                A(f"    ga[{i}] = fmaf({n}go[{j}], {mfx('b', k, i)}, ga[{i}]);")
    for k in range(B):
        first = True
        for i in range(B):
            j = out[i, k]
            n = "" if sign[i, k] > 0 else "-"
            if first:
                A(f"    gb[{k}] = {n}go[{j}] * {mfx('a', i, k)};")
                first = False
            else:
                A(f"    gb[{k}] = fmaf({n}go[{j}], {mfx('a', i, k)}, gb[{k}]);")
    A("  }")
    # ---- weighted geometric product, grouped by grade path
    # path p = (gi, gj, gk); members (i,k) with grade(i)=gi, grade(k)=gk, grade(out)=gj
    members = {}
    for i in range(B):
        for k in range(B):
            p = int(pidx[grade[i], grade[out[i, k]], grade[k]])
            members.setdefault(p, []).append((i, k))
    A("  // z[j] += sum over paths p of w[p] * (signed sum of x[i] r[k] over the (i,k) of p landing on j)")
    A("  template <bool MET> __device__ __forceinline__ static void wgp(const float* x, const float* r, const float* w, const float* mf, float* z) {")
    A("    float t;")
    for p in range(P):
        by_j = {}
        for (i, k) in members[p]:
            by_j.setdefault(int(out[i, k]), []).append((i, k))
        for j, lst in sorted(by_j.items()):
            first = True
            for (i, k) in lst:
                n = "" if sign[i, k] > 0 else "-"
                if first:
                    A(f"    t = {n}{mfx('x', i, k)} * r[{k}];")
                    first = False
                else:
                    A(f"    t = fmaf({n}{mfx('x', i, k)}, r[{k}], t);")
            A(f"    z[{j}] = fmaf(w[{p}], t, z[{j}]);")
    A("  }")
    A("  // adjoints: dx[i] += w[p] s dz[j] r[k]; dr[k] += w[p] s x[i] dz[j]; dw[p] += s x[i] r[k] dz[j]")
    A("  template <bool MET> __device__ __forceinline__ static void wgp_bwd(const float* x, const float* r, const float* w, const float* dz,")
    A("                                                 const float* mf, float* dx, float* dr, float* dw) {")
    A("    float t, u;")
    for p in range(P):
        by_i = {}
        by_k = {}
        for (i, k) in members[p]:
            by_i.setdefault(i, []).append(k)
            by_k.setdefault(k, []).append(i)
        A(f"    u = 0.f;  // path {p}")
        for i, ks in sorted(by_i.items()):
            first = True
            for k in ks:
                j = out[i, k]
                n = "" if sign[i, k] > 0 else "-"
                if first:
                    A(f"    t = {n}dz[{j}] * {mfx('r', k, i)};")
                    first = False
                else:
                    A(f"    t = fmaf({n}dz[{j}], {mfx('r', k, i)}, t);")
            A(f"    dx[{i}] = fmaf(w[{p}], t, dx[{i}]); u = fmaf(x[{i}], t, u);")
        A(f"    dw[{p}] += u;")
        for k, is_ in sorted(by_k.items()):
            first = True
            for i in is_:
                j = out[i, k]
                n = "" if sign[i, k] > 0 else "-"
                if first:
                    A(f"    t = {n}dz[{j}] * {mfx('x', i, k)};")
                    first = False
                else:
                    A(f"    t = fmaf({n}dz[{j}], {mfx('x', i, k)}, t);")
            A(f"    dr[{k}] = fmaf(w[{p}], t, dr[{k}]);")
    A("  }")
    A("};")
    A("")
    return "\n".join(L)


def main():
    parts = [
        "// GENERATED by gen_algebra.py from algebra/metric.py::product_table -- do not edit by hand.",
        "// Unrolled Euclidean Cl(n,0) product tables (reference: csmpn/algebra/metric.py:50-120,",
        "// csmpn/algebra/cliffordalgebra.py:44-54,238-252; csmpn/models/cegnn_utils.py:126-155).",
        "#pragma once",
        "",
        "namespace csmpn {",
        "template <int DIM_> struct Alg;",
        "",
    ]
    for d in DIMS:
        parts.append(emit_dim(d))
    parts.append("}  // namespace csmpn")
    path = os.path.join(HERE, "algebra_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(parts) + "\n")
    print("wrote", path, sum(p.count("\n") for p in parts), "lines")


if __name__ == "__main__":
    main()
