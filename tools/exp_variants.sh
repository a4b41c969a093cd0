#!/bin/bash
# run every experimental build of libamsweep through the same ticks (developer tool)
for lib in active-monitor_b200/lib/exp/libamsweep_*.so; do
  echo "== $lib config2"; AMSWEEP_LIB=$PWD/$lib python tools/prof_tick.py --ticks 6 | tail -3
  echo "== $lib config3"; AMSWEEP_LIB=$PWD/$lib python tools/prof_tick.py --config 3 --ticks 4 | tail -2
done
