#!/bin/bash
# round-2 GPU run #1: baseline of the round-1 code — e2e breakdown, compaction knobs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_run1.txt
nproc >> gpurun_out/r02_run1.txt
echo "== prof_e2e" >> gpurun_out/r02_run1.txt
timeout 300 python tools/prof_e2e.py >> gpurun_out/r02_run1.txt 2>&1
echo "== knobs" >> gpurun_out/r02_run1.txt
for lib in active-monitor_b200/lib/exp/libamsweep_g*_u*.so; do
  export AMSWEEP_LIB=$PWD/$lib
  echo -n "$lib: " >> gpurun_out/r02_run1.txt
  timeout 200 python bench.py --steps 200 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), 'G/s step', round(d['ms_per_step']*1e3,1), 'us  sweep', round(d['roofline']['kernel_ms']*1e3,1), 'us compact', round(d['roofline']['second_kernel']['kernel_ms']*1e3,1), 'e2e ms', round(d['e2e']['ms_per_step'],3))" >> gpurun_out/r02_run1.txt 2>&1
done
unset AMSWEEP_LIB
cat gpurun_out/r02_run1.txt
