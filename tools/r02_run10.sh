#!/bin/bash
# round-2 GPU run #10 (1 GPU): block kernel with warp-aggregated statistics and 128-thread CTAs; e2e with pinned consumer workers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run10
echo "== pytest -m gpu (1-GPU files)" > $O.txt
timeout 1200 python -m pytest tests/test_sweep_gpu.py tests/test_golden_fixtures.py tests/test_timezones.py -m gpu -x -q 2>&1 | tail -3 >> $O.txt
echo "== bench default" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
for w in 4 6; do timeout 300 python bench.py --steps 120 --warmup 10 --no-cpu --e2e-workers $w > $O.bench_w$w.json 2>> $O.txt; done
AMGEN_E2E_NO_PIN=1 timeout 300 python bench.py --steps 120 --warmup 10 --no-cpu --e2e-workers 10 > $O.bench_w10_nopin.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run10.bench.json"))
r=d["roofline"]
print("value G/s", round(d["value"]/1e9,2), "us/step", round(d["ms_per_step"]*1e3,2), "sweep us", round(r["kernel_ms"]*1e3,2), "rest us", round(r["rest_of_tick"]["ms"]*1e3,2), "frac", round(r["frac"],3), "step frac", round(r["step_level"]["frac"],3))
print("e2e default workers", d["e2e"]["consumer_workers"], round(d["e2e"]["ms_per_step"],4), d["e2e"]["split_ms_per_step"])
for w in ("w4","w6","w10_nopin"):
    e=json.load(open(f"gpurun_out/r02_run10.bench_{w}.json"))["e2e"]; print(w, round(e["ms_per_step"],4), e["split_ms_per_step"])
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","cores","ms_per_tick_min","ms_per_tick_median")})
PY
echo "== bench config 5" >> $O.txt
timeout 600 python bench.py --config 5 --steps 6000 --warmup 20 --no-cpu > $O.bench5.json 2>> $O.txt
python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run10.bench5.json"))
print("config5 us/tick", d["ms_per_step"]*1e3, "G/s", d["value"]/1e9, "blocked", d.get("temporal_blocking"))
PY
echo "== config 5, one simulated day, temporal blocking vs tick by tick" >> $O.txt
timeout 900 python tools/run_config5.py --blocked --compare-unblocked > $O.day_blocked.json 2>> $O.txt; cat $O.day_blocked.json >> $O.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_block_kernel" -s 2 -c 1 -o $O.blk -f python tools/run_config5.py --blocked --ticks 512 --sub 0 --full-ticks 0 > $O.ncu_blk.log 2>&1
tail -30 $O.txt
