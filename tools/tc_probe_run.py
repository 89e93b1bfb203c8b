#!/usr/bin/env python
"""[needs a diagnostics build: CSMPN_DEBUG_BUILD=1 python -c "import __graft_entry__ as g; g.build()"]
Run the tcgen05 probe (csmpn_tc_probe) over its modes and print the error of every variant against torch fp64.
Each variant runs in its own process so a faulting descriptor cannot poison the others.  GPU box only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(mode, M, N, K, flags, exact):
    import torch
    from csmpn_b200 import _lib
    from csmpn_b200._lib import check, lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(mode * 100 + N + K + flags)
    dev = torch.device("cuda:0")

    def rnd(*shape):
        x = torch.randn(*shape, generator=g)
        if exact:  # TF32-exact inputs: any correct configuration reproduces fp32 matmul to ~1e-6
            x = (x.view(torch.int32) & -8192).view(torch.float32)
        return x

    if mode == 0:
        A, B = rnd(128, K), rnd(N, K)
        ref = A.double() @ B.double().T
    elif mode == 1:
        A, B = rnd(128, K), rnd(K, N)
        ref = A.double() @ B.double()
    else:
        A, B = rnd(K, M), rnd(K, N)
        ref = A.double().T @ B.double()
    dump = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.to(dev), B.to(dev)  # keep the device copies alive across the launch
    check(lib().csmpn_tc_probe(mode, M, N, K, flags, ptr(Ad), ptr(Bd), ptr(dump), stream_ptr(dev)), "probe")
    torch.cuda.synchronize()
    d = dump.cpu().double()
    scale = float(ref.abs().max())
    out = {}
    if M == 128:
        out["rows=lanes"] = float((d - ref).abs().max() / scale)
    else:
        lanes = [(i % 16) + 32 * (i // 16) for i in range(64)]
        out["lane=(i%16)+32(i/16)"] = float((d[lanes] - ref).abs().max() / scale)
        out["lane=i"] = float((d[:64] - ref).abs().max() / scale)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        mode, M, N, K, flags, exact = map(int, sys.argv[1:7])
        print(one(mode, M, N, K, flags, bool(exact)))
        sys.exit(0)
    cases = []
    for flags in (0, 1, 2, 3):
        cases += [(0, 128, 32, 32, flags, 1), (1, 128, 32, 32, flags, 1), (2, 64, 32, 128, flags, 1), (2, 128, 32, 128, flags, 1)]
    cases += [(0, 128, 48, 40, 0, 1), (1, 128, 48, 40, 0, 1), (2, 64, 40, 128, 0, 1), (2, 64, 8, 64, 0, 1)]
    cases += [(0, 128, 32, 32, 0, 0), (0, 128, 32, 32, 4, 0), (1, 128, 32, 64, 4, 0), (2, 64, 32, 128, 4, 0)]
    for c in cases:
        r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True, timeout=300)
        tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else "FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1]
        print(f"mode={c[0]} M={c[1]} N={c[2]} K={c[3]} flags={c[4]} exact={c[5]}: {tail}", flush=True)
