#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/dist_tests.log 2>&1
echo "dist tests exit $?" | tee -a gpurun_out/dist_tests.log
grep -a "DIST_\|passed\|failed\|Error" gpurun_out/dist_tests.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
rm -f gpurun_out/bench_train.jsonl
CCDM_NO_OVERLAP=1 timeout 600 $TR --master-port 29721 tools/bench_train.py --model uk64 --graph --steps 10 2>&1 | grep '^{' 
timeout 600 $TR --master-port 29722 tools/bench_train.py --model uk64 --graph --steps 10 2>&1 | grep '^{'
timeout 600 python tools/bench_train.py --model uk64 --graph --steps 10 2>&1 | grep '^{'
