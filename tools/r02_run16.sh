#!/bin/bash
# round-2 GPU run #16 (1 GPU): temporal blocking A/B — 128-thread CTAs: ticks per block 64 / 128, register caps (min CTAs per SM 1 / 6 / 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run16.txt
: > $O
one() { # tag, lib, block ticks
  AMSWEEP_LIB=$2 AMSWEEP_BLOCK_TICKS=$3 timeout 300 python tools/run_config5.py --blocked --sub 0 --full-ticks 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'K=$3', round(d['us_per_tick_mean'],2), 'us/tick', round(d['device_ms_total']/1e3,3), 's/day', 'checksum', d['checksum_xor_of_idx_xor'], 'submits', d['total_submits'])" >> $O
}
L=$PWD/active-monitor_b200/lib
one default $L/libamsweep.so 64
one k128 $L/exp/libamsweep_k128.so 128
one min6 $L/exp/libamsweep_min6.so 64
one min8 $L/exp/libamsweep_min8.so 64
one min6_k128 $L/exp/libamsweep_min6_k128.so 128
one min6_k96 $L/exp/libamsweep_min6_k128.so 96
cat $O
