#!/bin/bash
# round-2 8-GPU run: parity at 4 and 8 ranks on hardware, standalone exchange, bench at N=8 (10 M and the
# literal 12.5 M per GPU of BASELINE configs[3]), round-1 plain push for comparison, N=4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02_mgpu_n8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
nvidia-smi topo -m > $O.topo.txt 2>&1
echo "== parity (tests/test_multi_gpu.py), worlds 4 and 8" > $O.txt
AMSWEEP_PUSH_TIMEOUT_MS=20000 timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "4- or 8-" 2>&1 | tail -6 >> $O.txt
echo "== standalone exchange, N=8" >> $O.txt
N=10000000 K=50 run 8 29801 tools/prof_gather.py 2>/dev/null | tail -1 > $O.exchange.json; cat $O.exchange.json >> $O.txt
echo "== bench N=8, 10 M per GPU" >> $O.txt
run 8 29802 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu 2>/dev/null | tail -1 > $O.bench.json; tail -c 1200 $O.bench.json >> $O.txt
echo "== bench N=8, 12.5 M per GPU (BASELINE configs[3] literal: 100 M records)" >> $O.txt
run 8 29803 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu --records-per-gpu 12500000 2>/dev/null | tail -1 > $O.bench_100m.json; tail -c 1200 $O.bench_100m.json >> $O.txt
echo "== bench N=8 plain (round-1 format)" >> $O.txt
run 8 29804 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu --gather plain --no-verify 2>/dev/null | tail -1 > $O.bench_plain.json; tail -c 500 $O.bench_plain.json >> $O.txt
echo "== bench N=4" >> $O.txt
run 4 29805 bench.py --gpus 4 --steps 200 --warmup 10 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r02_mgpu_n4.bench.json; tail -c 1200 gpurun_out/r02_mgpu_n4.bench.json >> $O.txt
python - <<'PY' >> $O.txt
import json
for f in ["gpurun_out/r02_mgpu_n8.bench.json","gpurun_out/r02_mgpu_n8.bench_100m.json","gpurun_out/r02_mgpu_n8.bench_plain.json","gpurun_out/r02_mgpu_n4.bench.json"]:
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], "GPUs", round(d["value"]/1e9,1), "G/s", round(d["ms_per_step"]*1e3,1), "us/step", "verified", d.get("gather_verified"), "e2e ms", round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, "failed", e)
PY
tail -30 $O.txt
