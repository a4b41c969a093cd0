#!/bin/bash
# round-2 GPU run #5 (1 GPU): full -m gpu suite on HEAD, bench lines (e2e with 1/4/8 consumer workers),
# e2e breakdown, config-3 timing experiment, ncu launch lists and full captures (config 2, config 3, e2e shape)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run5
echo "== pytest -m gpu (1-GPU files)" > $O.txt
timeout 1200 python -m pytest tests/test_sweep_gpu.py tests/test_golden_fixtures.py tests/test_timezones.py -m gpu -x -q 2>&1 | tail -3 >> $O.txt
echo "== smoke" >> $O.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $O.txt
echo "== bench default" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
for w in 1 2 8; do timeout 300 python bench.py --steps 120 --warmup 10 --no-cpu --e2e-workers $w > $O.bench_w$w.json 2>> $O.txt; done
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run5.bench.json"))
r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e", d["e2e"])
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","cores","ms_per_tick_min","ms_per_tick_median")})
for w in (1,2,8):
    try:
        e=json.load(open(f"gpurun_out/r02_run5.bench_w{w}.json"))["e2e"]; print("workers", w, round(e["ms_per_step"],4), e["split_ms_per_step"])
    except Exception as ex: print("workers", w, "failed", ex)
PY
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
echo "== exp_config3" >> $O.txt
timeout 300 python tools/exp_config3.py > $O.exp3.json 2>> $O.txt; cat $O.exp3.json >> $O.txt
echo "== bench config 3 / 5" >> $O.txt
timeout 600 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O.bench3.json 2>> $O.txt
timeout 600 python bench.py --config 5 --steps 6000 --warmup 20 --no-cpu > $O.bench5.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run5.bench3.json"))
print("config3 us/step", d["ms_per_step"]*1e3, "sweep", d["roofline"]["kernel_ms"]*1e3, "rest", d["roofline"]["rest_of_tick"]["ms"]*1e3, "frac", d["roofline"]["frac"], "clocks", d["clocks"])
d=json.load(open("gpurun_out/r02_run5.bench5.json"))
print("config5 us/tick", d["ms_per_step"]*1e3, "G/s", d["value"]/1e9)
PY
echo "== ncu launch lists" >> $O.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O.launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --settle-ms 5 > $O.ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O.launches3.csv python bench.py --config 3 --steps 3 --warmup 3 --no-cpu --settle-ms 5 > $O.ncu_bench3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O.launches5.csv python bench.py --config 5 --steps 30 --warmup 3 --no-cpu --settle-ms 1 > $O.ncu_bench5.log 2>&1
echo "== ncu full: config 2 tick, config 3 sweep, e2e-shaped tick" >> $O.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel|scan_groups" -s 8 -c 3 -o $O.c2 -f python tools/prof_tick.py --config 2 --ticks 5 > $O.ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel" -s 2 -c 1 -o $O.c3 -f python tools/prof_tick.py --config 3 --ticks 4 > $O.ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel|mark_ops|apply_result_ops|clear_marks" -s 25 -c 5 -o $O.e2e -f python tools/prof_tick.py --e2e --ticks 8 > $O.ncu_e2e.log 2>&1
ls -la gpurun_out/r02_run5* >> $O.txt
tail -40 $O.txt
