#!/usr/bin/env python
"""Hottest SASS regions of one kernel in an .ncu-rep: tools/ncu_hot.py rep kernel_regex [launch_index] [top]
Prints SASS instructions with the most stall samples, plus a coarse histogram (groups of 64 instructions)."""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
S = [(int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), k, r[ix["Source"]].strip()) for k, r in enumerate(data)]
tot = sum(s[0] for s in S); toti = sum(s[1] for s in S)
print(rows[0][1][:90], "samples", tot, "warp-instr", toti, "sass lines", len(S))
for s in sorted(S, reverse=True)[:top]:
    print(f"  {s[0]:6d} ({100*s[0]/max(tot,1):4.1f}%) exec={s[1]:8d} #{s[2]:5d} {s[3][:100]}")
print("-- histogram by 64-instruction group: samples%, instr%")
for g in range(0, len(S), 64):
    ss = sum(s[0] for s in S[g:g+64]); ii = sum(s[1] for s in S[g:g+64])
    if ss * 50 > tot or ii * 50 > toti:
        print(f"  #{g:5d}: {100*ss/max(tot,1):5.1f}% samples {100*ii/max(toti,1):5.1f}% instr   first: {S[g][3][:70]}")
