#!/usr/bin/env python
"""[needs a diagnostics build: CSMPN_DEBUG_BUILD=1 python -c "import __graft_entry__ as g; g.build()"]
Layout exploration with csmpn_tc_probe_raw: build shared-memory images for a hypothesis (layout function + descriptor
fields), run one tile on the tensor pipe, compare with torch.  One subprocess per experiment.  GPU box only."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NONE, SW128, SW64, SW32 = 0, 2, 4, 6


def lay_none(R):  # [C/4][R][4]
    return lambda r, c: ((c // 4) * R + r) * 16 + (c % 4) * 4


def lay_sw128(R):  # 32-channel blocks of [R][128 B], 16-byte chunk index xor (r % 8)
    return lambda r, c: (c // 32) * R * 128 + r * 128 + ((((c % 32) // 4) ^ (r % 8)) * 16) + (c % 4) * 4


def lay_sw32(R):  # 8-channel blocks of [R][32 B], chunk index xor ((r >> 2) & 1)
    return lambda r, c: (c // 8) * R * 32 + r * 32 + ((((c % 8) // 4) ^ ((r >> 2) & 1)) * 16) + (c % 4) * 4


def experiments():
    E = {}
    # name: (a_layout_fn name, a dims (R, C), a desc (mn, lbo, sbo, type, kinc), same for b, M, N, ksteps, kind)
    # K-major pairs: A [128 x K], B [N x K]
    K, N = 32, 32
    E["k_sw128"] = dict(kind="k", K=K, N=N, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("sw128", 0, 16, 1024, SW128, 32))
    E["k_none"] = dict(kind="k", K=K, N=N, M=128, a=("none", 0, 128 * 16, 128, NONE, 2 * 128 * 16), b=("none", 0, N * 16, 128, NONE, 2 * N * 16))
    E["k_none_swapped"] = dict(kind="k", K=K, N=N, M=128, a=("none", 0, 128, 128 * 16, NONE, 2 * 128 * 16), b=("none", 0, 128, N * 16, NONE, 2 * N * 16))
    E["k_sw32"] = dict(kind="k", K=K, N=N, M=128, a=("sw32", 0, 16, 256, SW32, 128 * 32), b=("sw32", 0, 16, 256, SW32, N * 32))
    E["k_sw32_lbo0"] = dict(kind="k", K=K, N=N, M=128, a=("sw32", 0, 0, 256, SW32, 128 * 32), b=("sw32", 0, 0, 256, SW32, N * 32))
    E["k_none_A_sw128_B"] = dict(kind="k", K=K, N=N, M=128, a=("none", 0, 128 * 16, 128, NONE, 2 * 128 * 16), b=("sw128", 0, 16, 1024, SW128, 32))
    E["k_sw128_A_none_B"] = dict(kind="k", K=K, N=N, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("none", 0, N * 16, 128, NONE, 2 * N * 16))
    E["k_sw128_A_none_B_swapped"] = dict(kind="k", K=K, N=N, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("none", 0, 128, N * 16, NONE, 2 * N * 16))
    # MN-major: A [K x M] (plane rows = K), B [K x N]
    Kr = 32
    for M in (64, 128):
        E[f"mn_none_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=("none", 1, 128, Kr * 16, NONE, 128), b=("none", 1, 128, Kr * 16, NONE, 128))
        E[f"mn_none_swapped_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=("none", 1, Kr * 16, 128, NONE, 128), b=("none", 1, Kr * 16, 128, NONE, 128))
        E[f"mn_sw128_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=("sw128", 1, Kr * 128, 1024, SW128, 1024), b=("sw128", 1, Kr * 128, 1024, SW128, 1024))
        E[f"mn_sw32_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=("sw32", 1, Kr * 32, 256, SW32, 256), b=("sw32", 1, Kr * 32, 256, SW32, 256))
    # mixed: A K-major sw128, B MN-major (transposed-weight GEMM)
    E["kA_sw128_mnB_sw128"] = dict(kind="kmn", K=32, N=32, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("sw128", 1, 32 * 128, 1024, SW128, 1024))
    E["kA_sw128_mnB_none"] = dict(kind="kmn", K=32, N=32, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("none", 1, 128, 32 * 16, NONE, 128))
    E["kA_sw128_mnB_none_swapped"] = dict(kind="kmn", K=32, N=32, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("none", 1, 32 * 16, 128, NONE, 128))
    E["kA_sw128_mnB_sw32"] = dict(kind="kmn", K=32, N=32, M=128, a=("sw128", 0, 16, 1024, SW128, 32), b=("sw32", 1, 32 * 32, 256, SW32, 256))
    for M in (64, 128):
        for lay in ("mn32b", "rowmajor"):
            E[f"mn_{lay}_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=(lay, 1, Kr * 128, 512, SW128_32B, 1024), b=(lay, 1, Kr * 128, 512, SW128_32B, 1024))
            E[f"mn_{lay}_swapped_M{M}"] = dict(kind="mn", K=Kr, N=32, M=M, a=(lay, 1, 512, Kr * 128, SW128_32B, 1024), b=(lay, 1, 512, Kr * 128, SW128_32B, 1024))
    E["mn_mn32b_M64_N48"] = dict(kind="mn", K=Kr, N=48, M=64, a=("mn32b", 1, Kr * 128, 512, SW128_32B, 1024), b=("mn32b", 1, Kr * 128, 512, SW128_32B, 1024))
    E["kA_none_mnB_mn32b"] = dict(kind="kmn", K=32, N=32, M=128, a=("none", 0, 128 * 16, 128, NONE, 2 * 128 * 16), b=("mn32b", 1, 32 * 128, 512, SW128_32B, 1024))
    E["decode_mn32b_A"] = dict(kind="decode", K=8, N=16, M=128, a=("none", 1, 8 * 128, 512, SW128_32B, 0), b=None)
    E["decode_mn32b_A_swapped"] = dict(kind="decode", K=8, N=16, M=128, a=("none", 1, 512, 8 * 128, SW128_32B, 0), b=None)
    # decode: B = ones, A word w holds value w; a single K step; prints the raw sums
    E["decode_none_A"] = dict(kind="decode", K=8, N=16, M=128, a=("none", 0, 128 * 16, 128, NONE, 0), b=None)
    E["decode_none_A_swapped"] = dict(kind="decode", K=8, N=16, M=128, a=("none", 0, 128, 128 * 16, NONE, 0), b=None)
    return E


def swz32b(w):  # SWIZZLE_128B_BASE32B: 4-byte word index inside a 16-byte chunk xor (chunk index & 3)
    return (w & ~3) | ((w & 3) ^ ((w >> 2) & 3))


def lay_mn32b(R):  # MN-major tf32 (SWIZZLE_128B_BASE32B): 32-channel groups of [R rows][128 B]; 32-byte unit index xor (r & 3)
    return lambda r, c: (c // 32) * R * 128 + r * 128 + ((((c % 32) // 8) ^ (r & 3)) * 32) + (c % 8) * 4


def lay_rowmajor(R):  # same without the intra-row swizzle (expects a fixed channel permutation of the result)
    return lambda r, c: (c // 32) * R * 128 + r * 128 + (c % 32) * 4


SW128_32B = 1
LAY = {"none": lay_none, "sw128": lay_sw128, "sw32": lay_sw32, "mn32b": lay_mn32b, "rowmajor": lay_rowmajor}


def run(name):
    import numpy as np
    import torch
    from csmpn_b200._lib import check, lib, ptr, stream_ptr

    e = experiments()[name]
    g = torch.Generator().manual_seed(1)
    dev = torch.device("cuda:0")
    K, N, M = e["K"], e["N"], e["M"]

    def rnd(*shape):
        x = torch.randn(*shape, generator=g)
        return (x.view(torch.int32) & -8192).view(torch.float32)

    def image(mat, layname, words):
        R, C = mat.shape
        fn = LAY[layname](R)
        img = np.zeros(words, dtype=np.float32)
        for r in range(R):
            for c in range(C):
                img[fn(r, c) // 4] = float(mat[r, c])
        return torch.from_numpy(img)

    if e["kind"] == "decode":
        a_words = 8192
        a_img = torch.arange(a_words, dtype=torch.float32)
        a_img[2048:] = 0.0
        b_img = torch.ones(4096)
        ref = None
    else:
        if e["kind"] == "k":
            A, B = rnd(128, K), rnd(N, K)
            ref = A.double() @ B.double().T
        elif e["kind"] == "mn":
            A, B = rnd(K, M), rnd(K, N)
            ref = A.double().T @ B.double()
        else:
            A, B = rnd(128, K), rnd(K, N)
            ref = A.double() @ B.double()
        a_img = image(A, e["a"][0], 16384)
        b_img = image(B, e["b"][0], 16384)
    a, b = e["a"], e["b"] or ("none", 0, 16, 16, NONE, 0)
    prm = (ctypes.c_uint32 * 16)(M, N, a[1], b[1], a[2], a[3], a[4], b[2], b[3], b[4], max(K // 8, 1), a[5], b[5], 0, 0, 0)
    dump = torch.full((128, N), float("nan"), device=dev)
    a_dev, b_dev = a_img.to(dev), b_img.to(dev)  # keep the device copies alive across the launch
    check(lib().csmpn_tc_probe_raw(ptr(a_dev), a_img.numel(), ptr(b_dev), b_img.numel(), prm, ptr(dump),
                                   stream_ptr(dev)), "probe_raw")
    torch.cuda.synchronize()
    d = dump.cpu().double()
    if ref is None:
        return {"D[0:40,0]": d[:40, 0].tolist(), "D[8::8,0]": d[8::8, 0].tolist(), "D[0,0:4]": d[0, :4].tolist()}
    scale = float(ref.abs().max())
    if e["a"][0] == "rowmajor":  # plain row-major planes: result channel m holds logical channel swz32b(m) of its 32-group
        pm = [(m // 32) * 32 + swz32b(m % 32) for m in range(ref.shape[0])]
        pn = [(n // 32) * 32 + swz32b(n % 32) for n in range(ref.shape[1])]
        ref = ref[pm][:, pn]
    if ref.shape[0] == 128:
        return {"err": float((d - ref).abs().max() / scale)}
    lanes = [(i % 16) + 32 * (i // 16) for i in range(64)]
    return {"err_lane16x4": float((d[lanes] - ref).abs().max() / scale), "err_lane_i": float((d[:64] - ref).abs().max() / scale)}


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(run(sys.argv[1])))
        sys.exit(0)
    only = os.environ.get("PROBE_ONLY", "")
    for name in experiments():
        if only and not any(tok in name for tok in only.split(",")):
            continue
        r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
        if r.returncode == 0:
            tail = (r.stdout.strip().splitlines() or [""])[-1]
        else:
            err = [l for l in r.stderr.strip().splitlines() if "rror" in l]
            tail = "FAILED: " + " | ".join(err[-2:])[:300]
        print(f"{name}: {tail}", flush=True)
