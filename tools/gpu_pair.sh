#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/pair_layers.log
for ms in 4 8; do
  echo "== CCDM_TAPGEMM_MINSTAGES=$ms" >> gpurun_out/pair_layers.log
  CCDM_TAPGEMM_MINSTAGES=$ms timeout 300 python tools/prof_layer.py conv64,conv64plain,conv128,res1x1 --iters 5 >> gpurun_out/pair_layers.log 2>&1
done
cat gpurun_out/pair_layers.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 8 gpurun_out/gpu_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
CCDM_TAPGEMM_MINSTAGES=8 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ms8.json 2> gpurun_out/bench.err
cut -c1-200 gpurun_out/bench_ms8.json; tail -n 3 gpurun_out/bench.err
