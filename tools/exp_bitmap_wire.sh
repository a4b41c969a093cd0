#!/bin/bash
# First hardware run of the experimental bitmap wire format (never executed on a GPU in round 1;
# logic-checked on the CPU emulator only).  Usage on a multi-GPU box:
#   gpurun --gpus 2 --timeout 420 -- 'bash tools/exp_bitmap_wire.sh 2'
#   gpurun --gpus 8 --timeout 300 -- 'bash tools/exp_bitmap_wire.sh 8'
# Every torchrun is wrapped in `timeout`: the push kernels spin on their peers.
N=${1:-2}
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== parity, world $N, all three wire formats"
AMSWEEP_TEST_EXPERIMENTAL_WIRES=1 timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -k "${N}-" 2>&1 | tail -3
p=29800
for w in plain c3 bm; do
  p=$((p+1))
  echo "== bench N=$N wire=$w"
  run $N $p bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --wire $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['n_gpus'], round(d['value']/1e9,1), 'G/s', round(d['ms_per_step']*1e3,1), 'us/step', d['gpu_launches'])"
done
