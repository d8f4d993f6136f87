#!/usr/bin/env bash
# Generic GPU pass: focused tests first, then the whole GPU suite, then the headline bench.
#   gpurun --timeout 1500 -- 'bash tools/gpu_pass.sh tests/test_gpu_linattn.py'
set -u
mkdir -p gpurun_out
FOCUS="${1:-}"
if [ -n "$FOCUS" ]; then
  timeout 600 python -m pytest $FOCUS -x -q -s > gpurun_out/focus_tests.log 2>&1
  echo "focus tests exit $?" | tee -a gpurun_out/focus_tests.log
fi
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -n 25 gpurun_out/focus_tests.log 2>/dev/null
tail -n 8 gpurun_out/gpu_tests.log
cat gpurun_out/bench.json
tail -n 5 gpurun_out/bench.err
