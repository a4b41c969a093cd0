#!/bin/bash
# round-2 multi-GPU run: parity of the tick exchange on hardware, bench at N, standalone exchange timing.
# usage: bash tools/r02_multigpu.sh <N> [records per GPU for the bench]
cd "$(dirname "$0")/.."
N=${1:-2}; NREC=${2:-10000000}
mkdir -p gpurun_out
O=gpurun_out/r02_mgpu_n$N
run() { timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
nvidia-smi topo -m > $O.topo.txt 2>&1
echo "== parity (tests/test_multi_gpu.py), worlds <= $N" > $O.txt
K=""; for w in 2 4 8; do [ $w -le $N ] && K="$K or ${w}-"; done; K="${K# or }"
AMSWEEP_PUSH_TIMEOUT_MS=20000 timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "$K" 2>&1 | tail -6 >> $O.txt
echo "== standalone exchange, N=$N" >> $O.txt
N=$NREC K=50 run $N 29801 tools/prof_gather.py > $O.exchange.json 2>> $O.txt; cat $O.exchange.json >> $O.txt
echo "== bench N=$N exchange" >> $O.txt
run $N 29802 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --records-per-gpu $NREC > $O.bench.json 2>> $O.txt; tail -c 2500 $O.bench.json >> $O.txt
echo "== bench N=$N plain (round-1 format)" >> $O.txt
run $N 29803 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --records-per-gpu $NREC --gather plain --no-verify > $O.bench_plain.json 2>> $O.txt; tail -c 600 $O.bench_plain.json >> $O.txt
tail -40 $O.txt
echo "== standalone exchange, push grid 16 / 592 CTAs" >> $O.txt
AMSWEEP_PUSH_CTAS=16 N=$NREC K=20 run $N 29804 tools/prof_gather.py 2>/dev/null | tail -1 | cut -c1-400 >> $O.txt
AMSWEEP_PUSH_CTAS=592 N=$NREC K=20 run $N 29805 tools/prof_gather.py 2>/dev/null | tail -1 | cut -c1-400 >> $O.txt
tail -12 $O.txt
