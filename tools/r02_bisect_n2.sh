#!/bin/bash
# which change cost the two-GPU step its overlap (111 -> 126 us)?  the same bench line from four historical trees and HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=$PWD/gpurun_out/r02_bisect_n2.txt
: > $O
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --no-verify; }
for c in 456cdb2 2021eba cde9c09 efecb20 HEAD; do
  if [ $c = HEAD ]; then d=.; else d=gpurun_wt/$c; fi
  (cd $d && run 29811 2>/dev/null | grep "^{" | tail -1 > /tmp/b.json; python - <<PY >> $O
import json
try:
    d=json.load(open("/tmp/b.json")); print("$c", round(d["ms_per_step"]*1e3,1), "us/step", round(d["value"]/1e9,1), "G/s sweep", round(d["roofline"]["kernel_ms"]*1e3,1))
except Exception as e: print("$c failed", e)
PY
)
done
cat $O
