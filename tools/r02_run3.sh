#!/bin/bash
# round-2 GPU run #3 (1 GPU): bench line, e2e breakdown, ncu captures of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run3
echo "== bench default" > $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
tail -c 3500 $O.bench.json >> $O.txt
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
echo "== bench config 3" >> $O.txt
timeout 600 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O.bench3.json 2>> $O.txt; python - <<'PY' >> $O.txt
import json
d=json.load(open("gpurun_out/r02_run3.bench3.json"))
print("config3 ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "rest", d["roofline"]["rest_of_tick"]["ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
PY
echo "== ncu launch lists" >> $O.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O.launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --settle-ms 5 > $O.ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O.launches5.csv python bench.py --config 5 --steps 30 --warmup 3 --no-cpu --settle-ms 1 > $O.ncu_bench5.log 2>&1
echo "== ncu full: config 2 tick (sweep, scan, expand, publish)" >> $O.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel|scan_groups" -s 8 -c 3 -o $O.c2 -f python tools/prof_tick.py --config 2 --ticks 5 > $O.ncu_c2.log 2>&1
echo "== ncu full: config 3 sweep" >> $O.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel|expand_kernel" -s 2 -c 2 -o $O.c3 -f python tools/prof_tick.py --config 3 --ticks 3 > $O.ncu_c3.log 2>&1
echo "== ncu full: config 5 closed loop off-minute" >> $O.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tick_kernel" -s 3 -c 1 -o $O.c5 -f python tools/prof_tick.py --config 5 --mode 1 --dt 1 --ticks 6 > $O.ncu_c5.log 2>&1
ls -la gpurun_out/ >> $O.txt
tail -40 $O.txt
