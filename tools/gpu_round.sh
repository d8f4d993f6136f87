#!/usr/bin/env bash
# Full GPU pass: test suite, bandwidth-kernel table, headline bench (with the training block).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 12 gpurun_out/gpu_tests.log
timeout 300 python tools/prof_bw.py > gpurun_out/prof_bw.log 2>&1; cat gpurun_out/prof_bw.log
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','e2e','train','cpu_baseline') if k in d}); print(d['roofline']['achieved'], d['roofline']['frac'])"; tail -n 3 gpurun_out/bench.err
