#!/usr/bin/env bash
# First GPU call of round 2: the vanilla (GroupNorm) UNet path that was built after the round-1 GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_pass.sh'
set -u
mkdir -p gpurun_out
CCDM_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_vanilla.py -q -s > gpurun_out/vanilla_tests.log 2>&1
echo "vanilla tests exit $?" | tee -a gpurun_out/vanilla_tests.log
timeout 600 python tools/bench_sample.py --model vrc64 --batch 200 --steps 1 --warmup 1 > gpurun_out/vanilla_bench.log 2>&1
echo "vanilla bench exit $?" | tee -a gpurun_out/vanilla_bench.log
timeout 600 python tools/bench_train.py --model vrc64 --batch 128 --steps 3 --warmup 2 --breakdown > gpurun_out/vanilla_train_bench.log 2>&1
echo "vanilla train bench exit $?" | tee -a gpurun_out/vanilla_train_bench.log
timeout 300 python tools/prof_groupnorm.py > gpurun_out/prof_groupnorm.log 2>&1
echo "prof_groupnorm exit $?" | tee -a gpurun_out/prof_groupnorm.log
tail -15 gpurun_out/vanilla_tests.log
tail -3 gpurun_out/vanilla_bench.log gpurun_out/vanilla_train_bench.log gpurun_out/prof_groupnorm.log
