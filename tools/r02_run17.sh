#!/bin/bash
# round-2 GPU run #17 (1 GPU): temporal blocking A/B, second round — register caps and ticks per block around the best of run 16
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run17.txt
: > $O
one() { # tag, lib, block ticks
  AMSWEEP_LIB=$2 AMSWEEP_BLOCK_TICKS=$3 timeout 300 python tools/run_config5.py --blocked --sub 0 --full-ticks 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'K=$3', round(d['us_per_tick_mean'],2), 'us/tick', round(d['device_ms_total']/1e3,3), 's/day', 'checksum', d['checksum_xor_of_idx_xor'], 'submits', d['total_submits'])" >> $O
}
L=$PWD/active-monitor_b200/lib
one default $L/libamsweep.so 64
one min8_k128 $L/exp/libamsweep_min8_k128.so 64
one min8_k128 $L/exp/libamsweep_min8_k128.so 96
one min8_k128 $L/exp/libamsweep_min8_k128.so 128
one min10 $L/exp/libamsweep_min10.so 64
one min12 $L/exp/libamsweep_min12.so 64
one t64_min16 $L/exp/libamsweep_t64_min16.so 64
one t256_min4 $L/exp/libamsweep_t256_min4.so 64
cat $O
