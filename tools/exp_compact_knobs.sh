#!/bin/bash
# Round-2 A/B of the compaction knobs on ONE GPU: builds variants of libamsweep.so into
# active-monitor_b200/lib/exp/ (git-ignored, travel with gpurun) — run the build part here on the
# CPU box first ("bash tools/exp_compact_knobs.sh build"), then on the GPU:
#   gpurun --timeout 600 -- 'bash tools/exp_compact_knobs.sh run'
# Hypothesis: 10 M records = 1221 compaction CTAs against 1184 resident slots,
# a nearly empty second wave of a latency-bound kernel.
cd "$(dirname "$0")/.."
if [ "$1" = "build" ]; then
  mkdir -p active-monitor_b200/lib/exp
  for v in "8 4" "16 4" "16 8" "32 8" "8 8"; do
    set -- $v
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
      -DAM_GROUP_TILES=$1 -DAM_COMPACT_UNROLL=$2 -o active-monitor_b200/lib/exp/libamsweep_g$1_u$2.so \
      active-monitor_b200/csrc/sweep.cu active-monitor_b200/csrc/gather.cu active-monitor_b200/csrc/cron_parse.cpp \
      active-monitor_b200/csrc/handoff.cpp && echo built g$1 u$2
  done
  exit 0
fi
for lib in active-monitor_b200/lib/exp/libamsweep_g*_u*.so; do
  export AMSWEEP_LIB=$PWD/$lib
  echo -n "$lib: "
  timeout 200 python bench.py --steps 200 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), 'G/s step', round(d['ms_per_step']*1e3,1), 'us  sweep', round(d['roofline']['kernel_ms']*1e3,1), 'us')"
  timeout 120 python -m pytest tests/test_sweep_gpu.py -m gpu -x -q -k "config2 or config3 or closed_loop" 2>&1 | tail -1
done
