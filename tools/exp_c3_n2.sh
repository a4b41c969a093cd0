#!/bin/bash
# 2-GPU validation of the compressed wire format: parity test, standalone exchange, bench
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 200 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -2
K=30 run 2 29651 tools/prof_gather.py 2>&1 | grep -E "^rank 0" | head -6
for w in plain c3; do
  run 2 29652 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --wire $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step', d['gpu_launches'])"
done
