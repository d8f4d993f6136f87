set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_linattn.py tests/test_gpu_backward.py tests/test_sngan.py -x -q > gpurun_out/focus_tests.log 2>&1
echo "focus tests exit $?" | tee -a gpurun_out/focus_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python tools/prof_linattn.py --iters 1 > gpurun_out/prof_linattn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:linattn_ -c 6 -o gpurun_out/prof_la python tools/prof_linattn.py --iters 1 > gpurun_out/ncu_la.log 2>&1
echo "ncu la exit $?"
tail -n 4 gpurun_out/focus_tests.log; cat gpurun_out/bench.json | cut -c1-300
