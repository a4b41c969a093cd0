#!/bin/bash
# round-2 GPU run #2 (1 GPU): first hardware run of the bitmap-native tick
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run2
echo "== smoke" > $O.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $O.txt 2>&1 || { echo "SMOKE FAILED" >> $O.txt; tail -30 $O.txt; exit 1; }
echo "== pytest -m gpu" >> $O.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 >> $O.txt
echo "== bench default" >> $O.txt
timeout 600 python bench.py --steps 200 --warmup 20 > $O.bench.json 2>> $O.txt
tail -c 3000 $O.bench.json >> $O.txt
echo "== prof_e2e" >> $O.txt
timeout 300 python tools/prof_e2e.py > $O.e2e.json 2>> $O.txt; cat $O.e2e.json >> $O.txt
echo "== bench config 3" >> $O.txt
timeout 600 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O.bench3.json 2>> $O.txt; tail -c 2500 $O.bench3.json >> $O.txt
echo "== bench config 5" >> $O.txt
timeout 600 python bench.py --config 5 --steps 6000 --warmup 20 --no-cpu > $O.bench5.json 2>> $O.txt; tail -c 2000 $O.bench5.json >> $O.txt
echo "== ncu launch list" >> $O.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O.launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --settle-ms 5 > $O.ncu_bench.log 2>&1
tail -5 $O.launches.csv >> $O.txt
tail -60 $O.txt
