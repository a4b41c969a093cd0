#!/bin/bash
# round-2 GPU run #18 (1 GPU): the state at the end of round 2 (after the e2e and temporal-blocking work) — tests, smoke, bench lines (both arms), breakdowns, ncu captures for profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run18
echo "== pytest -m gpu (1-GPU files)" > $O.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O.txt
echo "== smoke" >> $O.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $O.txt
echo "== bench default / reference arm" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O.bench_ref.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run18.bench.json"))
r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e workers", d["e2e"]["consumer_workers"], round(d["e2e"]["ms_per_step"],4), d["e2e"]["split_ms_per_step"])
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","cores","ms_per_tick_min","ms_per_tick_median")})
try:
    r=json.loads(open("gpurun_out/r02_run18.bench_ref.json").read().strip().splitlines()[-1]); print("reference arm", r.get("value"), r.get("ms_per_step"), r.get("cpu_baseline",{}).get("cores"))
except Exception as e: print("reference arm: parse failed", e)
PY
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
echo "== bench config 3 / 5" >> $O.txt
timeout 600 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O.bench3.json 2>> $O.txt
timeout 600 python bench.py --config 5 --steps 6000 --warmup 20 --no-cpu > $O.bench5.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run18.bench3.json"))
print("config3 us/step", d["ms_per_step"]*1e3, "sweep", d["roofline"]["kernel_ms"]*1e3, "rest", d["roofline"]["rest_of_tick"]["ms"]*1e3, "frac", d["roofline"]["frac"])
d=json.load(open("gpurun_out/r02_run18.bench5.json"))
print("config5 us/tick", d["ms_per_step"]*1e3, "G/s", d["value"]/1e9, "blocked", d.get("temporal_blocking"))
PY
echo "== config 5, one simulated day, temporal blocking vs tick by tick" >> $O.txt
timeout 900 python tools/run_config5.py --blocked --compare-unblocked > $O.day_blocked.json 2>> $O.txt; cat $O.day_blocked.json >> $O.txt
echo "== ncu" >> $O.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O.launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --settle-ms 5 > $O.ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel|scan_groups" -s 8 -c 3 -o $O.c2 -f python tools/prof_tick.py --config 2 --ticks 5 > $O.ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel" -s 2 -c 1 -o $O.c3 -f python tools/prof_tick.py --config 3 --ticks 4 > $O.ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel|mark_ops|apply_result|clear_marks" -s 30 -c 6 -o $O.e2e -f python tools/prof_tick.py --e2e --ticks 8 > $O.ncu_e2e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_block_kernel" -s 2 -c 1 -o $O.blk -f python tools/run_config5.py --blocked --ticks 512 --sub 0 --full-ticks 0 > $O.ncu_blk.log 2>&1
ls -la gpurun_out/r02_run18* >> $O.txt
tail -40 $O.txt
