#!/usr/bin/env bash
# tests, then the headline bench with and without an A/B switch given as $1 (e.g. CCDM_RESACC=0)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 6 gpurun_out/gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench.json').read()); print('A', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"; tail -n 3 gpurun_out/bench.err
cp gpurun_out/bench_layers.json gpurun_out/bench_layers_A.json
env $1 timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/bench_B.json 2> gpurun_out/bench.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_B.json').read()); print('B $1', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"; tail -n 3 gpurun_out/bench.err
cp gpurun_out/bench_layers.json gpurun_out/bench_layers_B.json
