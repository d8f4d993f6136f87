#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/bench_train.py --model uk64 --steps 3 --warmup 2 --breakdown > gpurun_out/train_eager.log 2>&1
tail -n 2 gpurun_out/train_eager.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/train_launches_uk64.csv python tools/bench_train.py --model uk64 --steps 1 --warmup 1 > gpurun_out/ncu_train.log 2>&1
echo "ncu exit $?"
python tools/ncu_summary.py gpurun_out/train_launches_uk64.csv gpurun_out/train_launches_uk64_summary.csv "UK64 batch 128 eager training step x2 (warm-up + timed), r2" > /dev/null 2>&1
head -n 45 gpurun_out/train_launches_uk64_summary.csv
