#!/bin/bash
# bench only at N GPUs (exchange), and the standalone exchange: usage r02_multigpu_bench.sh <N> [tag]
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-b}
mkdir -p gpurun_out
O=gpurun_out/r02_mgpu_n${N}_$TAG
run() { timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run $N 29802 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu 2>/dev/null | grep "^{" | tail -1 > $O.bench.json
N=10000000 K=50 run $N 29801 tools/prof_gather.py 2>/dev/null | grep "^{" | tail -1 > $O.exchange.json
python - <<PY
import json
d=json.load(open("$O.bench.json")); print("N=$N", round(d["value"]/1e9,1), "G/s", round(d["ms_per_step"]*1e3,1), "us/step host issue", round(d.get("host_issue_ms_per_step",0)*1e3,1), "verified", d.get("gather_verified"), "sweep", round(d["roofline"]["kernel_ms"]*1e3,1), "e2e ms", round(d["e2e"]["ms_per_step"],3))
e=json.load(open("$O.exchange.json")); print("exchange", e["us_per_exchange_max_over_ranks"], e["rank0_us_push_rebuild_publish_after_a_barrier"])
PY
