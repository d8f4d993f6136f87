#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/dist_tests.log 2>&1
echo "dist tests exit $?" | tee -a gpurun_out/dist_tests.log
grep -a "DIST_\|passed\|failed\|Error" gpurun_out/dist_tests.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['train'])"; tail -n 3 gpurun_out/bench_n2.err
