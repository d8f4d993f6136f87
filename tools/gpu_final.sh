#!/usr/bin/env bash
# Closing verification: smoke(), the GPU suite twice (flakiness), the full default bench line, the reference arm.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
for i in 1 2; do
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$i.log 2>&1
  echo "gpu tests run $i exit $?"; tail -n 2 gpurun_out/gpu_tests_$i.log
done
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python -c "import json; d=json.loads(open('gpurun_out/bench_final.json').read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['train']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"; tail -n 2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cut -c1-300 gpurun_out/bench_reference.json
