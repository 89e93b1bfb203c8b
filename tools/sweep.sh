#!/bin/bash
# BASELINE.json configs[4]: per-layer fwd+bwd sweep over hidden width 32-256 and 1 M - 100 M simplices (processed as
# successive independent launches of <= ~1 M simplices, inputs resident).  Writes one JSON line per configuration.
out=${1:-gpurun_out/r02_sweep.jsonl}
: > $out
run() { timeout 900 python bench.py --only --no-train --no-cpu-baseline --no-e2e "$@" >> $out 2>> ${out%.jsonl}.err || echo "{\"failed\": \"$*\"}" >> $out; }
run --hidden 64
run --hidden 128
run --hidden 256 --steps 10
run --hidden 64 --complexes 4000 --steps 5 --warmup 3
run --hidden 128 --complexes 1500 --steps 5 --warmup 3
run --hidden 256 --complexes 500 --steps 5 --warmup 3
run --complexes 11500 --steps 10 --warmup 3 --no-roofline          # C = 32: 1 M simplices per launch, 10 M in the timed region
run --complexes 11500 --steps 100 --warmup 3 --no-roofline         # 100 M
run --hidden 64 --complexes 4000 --steps 29 --warmup 3 --no-roofline   # C = 64: 10 M
