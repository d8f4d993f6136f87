#!/usr/bin/env bash
# BASELINE.json configuration matrix on N GPUs of one box (N = number of visible GPUs): appends to gpurun_out/*.jsonl
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
rm -f gpurun_out/bench_sample.jsonl gpurun_out/bench_train.jsonl
# headline + training block at N GPUs (what the driver's SCALE run records)
timeout 900 $TR --master-port 29801 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('bench', d['n_gpus'], d['value'], d['e2e']['value']); print(d.get('train'))"
# configs[2]: UK128 CcDPM DDPM-1000, batch 256 sharded over the box
timeout 900 $TR --master-port 29802 tools/bench_sample.py --model uk128 --global-batch 256 --steps 1 --warmup 1 2>&1 | grep '^{'
# configs[3]: SA128 DDIM-250 with covariance-embedded labels: weak (64 per GPU) and strong (64 total) scaling
timeout 900 $TR --master-port 29803 tools/bench_sample.py --model sa128 --batch 64 --steps 1 --warmup 1 2>&1 | grep '^{'
timeout 900 $TR --master-port 29804 tools/bench_sample.py --model sa128 --global-batch 64 --steps 2 --warmup 1 2>&1 | grep '^{'
# configs[4]: UK192 data-parallel training, 16 per GPU x 4 accumulated micro-batches, whole step as one CUDA graph
timeout 900 $TR --master-port 29805 tools/bench_train.py --model uk192 --batch 16 --accum 4 --graph --steps 5 2>&1 | grep '^{'
# configs[1]: UK64 training, batch 128 per GPU
timeout 900 $TR --master-port 29806 tools/bench_train.py --model uk64 --batch 128 --graph --steps 10 2>&1 | grep '^{'
