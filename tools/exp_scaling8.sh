#!/bin/bash
# 8-GPU experiment: push grid-size sweep (standalone), then bench at N=8 / N=4
run() { timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for c in 148 296 592 1184; do
  echo "== standalone push, 8 ranks, CTAS=$c"
  AMSWEEP_PUSH_CTAS=$c K=30 run 8 $((29600 + c % 97)) tools/prof_gather.py 2>&1 | grep -E "^rank 0" | head -3
done
for c in 296 1184; do
  echo "== bench N=8 CTAS=$c"
  AMSWEEP_PUSH_CTAS=$c run 8 $((29700 + c % 97)) bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step')"
done
echo "== bench N=4 CTAS=296"
AMSWEEP_PUSH_CTAS=296 run 4 29811 bench.py --gpus 4 --steps 200 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step')"
echo "== bench N=2 default"
run 2 29812 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step')"
