#!/usr/bin/env bash
# Round-2 ncu evidence: (1) launch list of one guided forward with durations + DRAM bytes (shares, traffic);
# (2) --set full of the isolated conv layers (CTA-pair 128->64, single-CTA 64->64) with source, for tensor-pipe / stall numbers.
set -u
mkdir -p gpurun_out
python tools/prof_forward.py --iters 2 > gpurun_out/prof_forward.log 2>&1 || { cat gpurun_out/prof_forward.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_fwd_r2.csv python tools/prof_forward.py --iters 2 > gpurun_out/ncu_fwd.log 2>&1
echo "ncu launch list exit $?"
python tools/prof_layer.py conv64,conv128 --iters 3 > gpurun_out/prof_layer.log 2>&1 && cat gpurun_out/prof_layer.log
ncu --set full --clock-control none --import-source on -k regex:tapgemm -c 4 -o gpurun_out/prof_tg_r2 -f \
    python tools/prof_layer.py conv64,conv128 --iters 1 > gpurun_out/ncu_tg.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/prof_tg_r2.ncu-rep
