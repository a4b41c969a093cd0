#!/bin/bash
# round-2 GPU run #22 (1 GPU): bench.py bound to the GPU's NUMA node — e2e spread over two invocations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run22
: > $O.txt
for r in a b; do
  timeout 600 python bench.py --steps 200 --warmup 20 $( [ $r = b ] && echo --no-cpu ) > $O.bench_$r.json 2>> $O.txt
  python - <<PY >> $O.txt
import json
d=json.load(open("$O.bench_$r.json")); r=d["roofline"]
print("$r value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"],4), d["e2e"]["runs_ms_per_step"], d["e2e"]["host_cpus"], d["e2e"]["split_ms_per_step"], "cpu", (d["cpu_baseline"] or {}).get("value"), (d["cpu_baseline"] or {}).get("cores"))
PY
done
cat $O.txt
