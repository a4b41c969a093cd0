#!/bin/bash
# round-2 GPU run #23 (1 GPU): e2e consumer workers 4..16 with the process bound to the GPU's NUMA node
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run24
: > $O.txt
for w in 4 10 16 10; do
  timeout 300 python bench.py --steps 120 --warmup 10 --no-cpu --e2e-workers $w > $O.bench_w$w.json 2>> $O.txt
  python - <<PY >> $O.txt
import json
d=json.load(open("$O.bench_w$w.json"))
print("workers $w e2e", round(d["e2e"]["ms_per_step"],4), d["e2e"]["runs_ms_per_step"], d["e2e"]["split_ms_per_step"])
PY
done
cat $O.txt
