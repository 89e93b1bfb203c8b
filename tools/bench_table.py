#!/usr/bin/env python
"""Markdown rows for DESIGN.md from a bench.py JSON line: tools/bench_table.py profiles/r01_bench_md17_default.json"""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {j['ms_per_step']:.3f} ms = {j['value']/1e6:.2f} M {j['unit']}; e2e {j['e2e']['ms_per_step']:.3f} ms = {j['e2e']['value']/1e6:.2f} M; "
      f"launches/step {j.get('gpu_launches_per_step')}; clocks {j['clocks']}")
if j.get("train"):
    print(f"train {j['train']['ms_per_step']:.2f} ms = {j['train']['value']/1e3:.2f} k complexes/s")
if j.get("cpu_baseline"):
    print("cpu", j["cpu_baseline"])
r = j["roofline"]
print("dominant:", r["kernel"], f"{r['frac']:.3f}" if r.get("frac") else None)
print("| kernel | us | algorithmic MB | DRAM MB (ncu) | GB/s | of peak |\n|---|---|---|---|---|---|")
for k in r.get("kernels", []):
    t = k.get("traffic")
    print(f"| {k['kernel']} | {k['launch_ms']*1e3:.0f} | {k['algorithmic_bytes']/1e6:.0f} | {t/1e6:.0f} | {k['hbm_gbs']:.0f} | {100*k['hbm_frac']:.0f} % |" if t else
          f"| {k['kernel']} | {k['launch_ms']*1e3:.0f} | {k['algorithmic_bytes']/1e6:.0f} | - | {k['hbm_gbs']:.0f} | {100*k['hbm_frac']:.0f} % |")
