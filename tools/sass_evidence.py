#!/usr/bin/env python
"""profiles/rNN_sass_tensor_core_kernels.txt: per tensor-core kernel of libcsmpn_b200.so the counts of the Blackwell-specific
SASS mnemonics (cuobjdump -sass) and a few sample lines.  usage: sass_evidence.py <out.txt> [round tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clifford-group-equivariant-simplicial-message-passing-networks_b200", "csrc", "libcsmpn_b200.so")
KEYS = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "ENL2.256", "UTCATOMSWS", "FFMA", "MUFU")


def main():
    out = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "round 2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            funcs[cur].append(line.strip())
    with open(out, "w") as f:
        f.write(f"# SASS evidence for the tensor-core kernels of libcsmpn_b200.so (cuobjdump -sass, sm_100a), {tag}\n"
                "# per kernel: counts of the Blackwell-specific mnemonics -- UTCHMMA = tcgen05.mma (kind::tf32), LDTM / STTM = tcgen05.ld / st (TMEM),\n"
                "# UBLKCP = cp.async.bulk (bulk copy engine, mbarrier completion), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, STG.E.ENL2.256 = 256-bit stores\n\n")
        for name, lines in funcs.items():
            if "3tcb" not in name or not any("UTCHMMA" in l for l in lines):
                continue
            cnt = {k: sum(k in l for l in lines) for k in KEYS}
            f.write(name + "\n")
            f.write(f"    instructions {len(lines)}: " + ", ".join(f"{k} {v}" for k, v in cnt.items() if v) + "\n")
            shown = 0
            for l in lines:
                if ("UBLKCP" in l or "UTCHMMA" in l) and shown < 4:
                    f.write("      " + re.sub(r"\s+", " ", l) + "\n")
                    shown += 1
    print("wrote", out)


if __name__ == "__main__":
    main()
