#!/usr/bin/env python
"""Recipe for oracle/_ref/: a verbatim, UNMODIFIED copy of the reference's Python sources (csmpn/ and engineer/, *.py
only) taken from /root/reference where they lie.  TEST / BASELINE INFRASTRUCTURE.

The reference is pure Python, so "building" it is copying it.  oracle/_ref/ is git-ignored (reference sources never
enter this repository's history) but travels to the GPU box with the working tree, where /root/reference does not
exist; there `bench.py --impl reference` (cpu_baseline.kind = "reference") and the optional reference-backed tests
import it through oracle/refshim.py, which supplies stand-ins for the third-party packages the reference imports
(torch_geometric, torch_scatter, gudhi).  __graft_entry__.build() runs this whenever /root/reference is present.

    python oracle/make_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("csmpn", "engineer")


def make_ref(src="/root/reference", dst=DST, quiet=False):
    if not os.path.isdir(os.path.join(src, "csmpn")):
        return None
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(src, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith((".py", ".yaml")):
                    continue
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    with open(os.path.join(dst, "README"), "w") as f:
        f.write("Verbatim copy of the reference's Python sources made by oracle/make_ref.py; not tracked in git.\n")
    if not quiet:
        print(f"copied {n} files from {src} to {dst}")
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    if make_ref(a.src) is None:
        raise SystemExit(f"no reference at {a.src}")
