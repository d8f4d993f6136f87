#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 4 gpurun_out/gpu_tests.log
rm -f gpurun_out/bench_train.jsonl
timeout 600 python tools/bench_train.py --model uk64 --graph --steps 10 2>&1 | grep '^{'
timeout 600 python tools/bench_train.py --model rc64 --graph --steps 10 2>&1 | grep '^{'

