#!/bin/bash
# 8-GPU comparison of the two wire formats: parity at world 8 and 4, bench at N=8 and N=4
run() { timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 150 python -m pytest tests/test_multi_gpu.py -x -q -k "8-c3 or 4-c3 or 8-plain" 2>&1 | tail -2
p=29700
for n in 8 4; do for w in c3 plain; do
  p=$((p+1))
  run $n $p bench.py --gpus $n --steps 200 --warmup 10 --no-cpu --wire $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step', d['gpu_launches'])"
done; done
