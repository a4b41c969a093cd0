#!/bin/bash
# run every experimental build of libamsweep through the same ticks (developer tool)
for lib in active-monitor_b200/lib/exp/libamsweep_*.so; do
  export AMSWEEP_LIB=$PWD/$lib
  echo "== $lib"
  echo -n "on-minute : "; timeout 60 python tools/prof_tick.py --ticks 6 2>&1 | tail -1
  echo -n "off-minute: "; timeout 60 python tools/prof_tick.py --config 2 --dt 1 --ticks 6 2>&1 | tail -1
  echo -n "closed    : "; timeout 60 python tools/prof_tick.py --config 5 --mode 1 --dt 1 --ticks 8 2>&1 | tail -1
  echo -n "config3   : "; timeout 60 python tools/prof_tick.py --config 3 --ticks 4 2>&1 | tail -1
  echo -n "bench     : "; timeout 120 python bench.py --steps 200 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), 'G/s step', round(d['ms_per_step']*1e3,1), 'us sweep', round(d['roofline']['kernel_ms']*1e3,1), 'us frac', round(d['roofline']['frac'],3))"
done
