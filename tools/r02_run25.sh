#!/bin/bash
# round-2 GPU run #25 (1 GPU): the final tree — all GPU tests, smoke, both bench arms, configs 3 and 5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run25
echo "== pytest -m gpu (all)" > $O.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O.txt
echo "== smoke" >> $O.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $O.txt
echo "== bench default / reference arm / config 3 / config 5" >> $O.txt
timeout 600 python bench.py > $O.bench.json 2>> $O.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O.bench_ref.json 2>> $O.txt
timeout 600 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O.bench3.json 2>> $O.txt
timeout 600 python bench.py --config 5 --steps 6000 --warmup 20 --no-cpu > $O.bench5.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run25.bench.json")); r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e", d["e2e"]["consumer_workers"], round(d["e2e"]["ms_per_step"],4), d["e2e"]["runs_ms_per_step"], d["e2e"]["host_cpus"], d["e2e"]["split_ms_per_step"])
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","cores","ms_per_tick_min","ms_per_tick_median")})
r=json.loads(open("gpurun_out/r02_run25.bench_ref.json").read().strip().splitlines()[-1]); print("reference arm", r.get("value"), r.get("ms_per_step"))
d=json.load(open("gpurun_out/r02_run25.bench3.json")); print("config3 us/step", d["ms_per_step"]*1e3, "sweep", d["roofline"]["kernel_ms"]*1e3, "frac", d["roofline"]["frac"])
d=json.load(open("gpurun_out/r02_run25.bench5.json")); print("config5 us/tick", d["ms_per_step"]*1e3, "blocked", d.get("temporal_blocking"))
PY
tail -12 $O.txt
