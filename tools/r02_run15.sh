#!/bin/bash
# round-2 GPU run #15 (1 GPU): temporal blocking A/B — CTA size of sweep_block_kernel (256 / 128 / 64 threads), ticks per block (64 / 32 / 16)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_run15.txt
: > $O
one() { # tag, lib, block ticks
  AMSWEEP_LIB=$2 AMSWEEP_BLOCK_TICKS=$3 timeout 300 python tools/run_config5.py --blocked --sub 0 --full-ticks 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'K=$3', round(d['us_per_tick_mean'],2), 'us/tick', round(d['device_ms_total']/1e3,3), 's/day', 'checksum', d['checksum_xor_of_idx_xor'], 'submits', d['total_submits'])" >> $O
}
L=$PWD/active-monitor_b200/lib
one t256 $L/libamsweep.so 64
one t128 $L/exp/libamsweep_blk128.so 64
one t64 $L/exp/libamsweep_blk64.so 64
one t256 $L/libamsweep.so 32
one t128 $L/exp/libamsweep_blk128.so 32
one t256 $L/libamsweep.so 16
cat $O
