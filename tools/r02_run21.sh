#!/bin/bash
# round-2 GPU run #21 (1 GPU): interleaved staged ops + fused drain-time result kernel: tests, bench, e2e breakdown, ncu of the drain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run21
echo "== pytest -m gpu (all)" > $O.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O.txt
echo "== bench default" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run21.bench.json"))
r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e workers", d["e2e"]["consumer_workers"], round(d["e2e"]["ms_per_step"],4), d["e2e"]["split_ms_per_step"], d["e2e"]["runs_ms_per_step"])
PY
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mark_ops|apply_result|clear_marks" -s 12 -c 3 -o $O.e2e -f python tools/prof_tick.py --e2e --ticks 8 > $O.ncu_e2e.log 2>&1
tail -12 $O.txt
