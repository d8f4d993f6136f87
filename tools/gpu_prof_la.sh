set -u
mkdir -p gpurun_out
python tools/prof_linattn.py > gpurun_out/prof_linattn.log 2>&1 && python tools/prof_linattn.py --n 1024 --C 128 >> gpurun_out/prof_linattn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:linattn_ -s 3 -c 3 -o gpurun_out/prof_la python tools/prof_linattn.py --iters 1 > gpurun_out/ncu_la.log 2>&1
echo "ncu la exit $?"
python tools/prof_forward.py > gpurun_out/prof_forward.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_fwd.csv python tools/prof_forward.py > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd exit $?"
cat gpurun_out/prof_linattn.log
