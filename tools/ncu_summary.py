#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel duration, issue utilisation, DRAM traffic, top stall reasons, hottest SASS lines."""
import collections
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_sector_hit_rate.pct"]
    for r in raw[2:]:
        if len(r) != len(hdr):
            continue
        print("====", r[idx["Kernel Name"]][:70], "grid", r[idx["Grid Size"]] if "Grid Size" in idx else "")
        for w in want:
            if w in idx:
                print(f"   {w}: {r[idx[w]]} {raw[1][idx[w]]}")
        st = []
        for h in stall:
            try:
                st.append((float(r[idx[h]].replace(",", "")), h))
            except ValueError:
                pass
        for v, h in sorted(st, reverse=True)[:6]:
            print("      stall", round(v, 2), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
    if topn <= 0:
        return
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    cur, hdr2 = None, None
    for r in src:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = r[1][:70]
            hdr2 = None
            rows = []
            continue
        if "# Samples" in r:
            hdr2 = {h: i for i, h in enumerate(r)}
            data = []
            name = cur
            continue
        if hdr2 and len(r) >= len(hdr2):
            data.append(r)
        elif hdr2 and data:
            pass
    # simple second pass: split by kernel
    blocks, cur = [], None
    for r in src:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1][:70], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and "# Samples" in r:
            cur["hdr"] = {h: i for i, h in enumerate(r)}
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for b in blocks:
        h = b["hdr"]

        def I(r, k):
            try:
                return int(float(r[h[k]]))
            except (ValueError, KeyError):
                return 0

        tot = sum(I(r, "# Samples") for r in b["rows"]) or 1
        toti = sum(I(r, "Instructions Executed") for r in b["rows"]) or 1
        print("---- hot SASS of", b["name"], "samples", tot, "inst", toti)
        byop = collections.Counter()
        for r in b["rows"]:
            t = r[h["Source"]].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            byop[op] += I(r, "Instructions Executed")
        print("   opcode mix:", ", ".join(f"{o} {100*c/toti:.0f}%" for o, c in byop.most_common(10)))
        for r in sorted(b["rows"], key=lambda r: -I(r, "# Samples"))[:topn]:
            print(f"   {100*I(r,'# Samples')/tot:5.1f}% x{I(r,'Instructions Executed'):8d} {r[h['Source']].strip()[:70]:70s} long_sb={r[h['stall_long_sb']]} bar={r[h['stall_barrier']]} short={r[h['stall_short_sb']]} wait={r[h['stall_wait']]} lg={r[h['stall_lg']]}")


if __name__ == "__main__":
    main()
