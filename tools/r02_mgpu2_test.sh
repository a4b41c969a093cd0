#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AMSWEEP_PUSH_TIMEOUT_MS=20000 timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "2-" 2>&1 | tail -6 | tee gpurun_out/r02_mgpu_n2_final_tests.txt
