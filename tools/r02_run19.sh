#!/bin/bash
# round-2 GPU run #19 (1 GPU): e2e with even pieces per worker and the deferred marks clear
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run19
echo "== pytest -m gpu (1-GPU files)" > $O.txt
timeout 1200 python -m pytest tests/test_sweep_gpu.py tests/test_golden_fixtures.py tests/test_timezones.py -m gpu -q 2>&1 | tail -3 >> $O.txt
echo "== bench default" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run19.bench.json"))
r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e workers", d["e2e"]["consumer_workers"], round(d["e2e"]["ms_per_step"],4), d["e2e"]["split_ms_per_step"])
PY
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
tail -12 $O.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_run19.launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --settle-ms 5 > gpurun_out/r02_run19.ncu_bench.log 2>&1
