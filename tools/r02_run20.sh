#!/bin/bash
# round-2 GPU run #20 (1 GPU): confirmation of the final tree + A/B: launch_dependents at the start of the sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run20
echo "== pytest -m gpu (all)" > $O.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O.txt
for v in default trig default trig; do
  if [ $v = default ]; then L=$PWD/active-monitor_b200/lib/libamsweep.so; else L=$PWD/active-monitor_b200/lib/exp/libamsweep_trig.so; fi
  AMSWEEP_LIB=$L timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu > $O.bench_$v.json 2>> $O.txt
  python - <<PY >> $O.txt
import json
d=json.load(open("$O.bench_$v.json")); r=d["roofline"]
print("$v", "value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"],4))
PY
done
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
tail -8 $O.txt
