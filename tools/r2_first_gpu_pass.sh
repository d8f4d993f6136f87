#!/usr/bin/env bash
# First GPU call of the next round: everything that was built after the round-1 GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_pass.sh'
# 1. the opt-in parity tests of the vanilla (GroupNorm) UNet path (all 62 pass in a host dry run, CCDM_GPU_TESTS_ON_HOST=1)   2. its DDIM-250 throughput at the RC-49 script config
# 3. HBM roofline of the GroupNorm row kernels                       4. the verified suite + headline bench, unchanged
set -u
mkdir -p gpurun_out
CCDM_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_vanilla.py -x -q -s > gpurun_out/vanilla_tests.log 2>&1
echo "vanilla tests exit $?" | tee -a gpurun_out/vanilla_tests.log
timeout 600 python tools/bench_sample.py --model vrc64 --batch 200 --steps 1 --warmup 1 > gpurun_out/vanilla_bench.log 2>&1
echo "vanilla bench exit $?" | tee -a gpurun_out/vanilla_bench.log
timeout 600 python tools/bench_train.py --model vrc64 --batch 128 --steps 3 --warmup 2 --breakdown > gpurun_out/vanilla_train_bench.log 2>&1
echo "vanilla train bench exit $?" | tee -a gpurun_out/vanilla_train_bench.log
timeout 300 python tools/prof_groupnorm.py > gpurun_out/prof_groupnorm.log 2>&1
echo "prof_groupnorm exit $?" | tee -a gpurun_out/prof_groupnorm.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err
tail -3 gpurun_out/vanilla_tests.log gpurun_out/vanilla_bench.log gpurun_out/vanilla_train_bench.log gpurun_out/prof_groupnorm.log gpurun_out/gpu_tests.log
cat gpurun_out/bench_r2_first.json
