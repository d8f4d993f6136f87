#!/usr/bin/env bash
# Quick GPU pass: focused tests, isolated linear-attention timings, headline bench (no CPU baseline).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest ${1:-tests/test_gpu_linattn.py} -x -q > gpurun_out/focus_tests.log 2>&1
echo "focus tests exit $?" | tee -a gpurun_out/focus_tests.log
python tools/prof_linattn.py > gpurun_out/prof_linattn.log 2>&1; python tools/prof_linattn.py --n 1024 --C 128 >> gpurun_out/prof_linattn.log 2>&1
python tools/prof_linattn.py --n 1024 --C 64 >> gpurun_out/prof_linattn.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 6 gpurun_out/focus_tests.log; cat gpurun_out/prof_linattn.log; cat gpurun_out/bench.json | cut -c1-400; tail -n 3 gpurun_out/bench.err
