#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/prof_layer.py conv64,conv64plain,conv128 --halo --iters 5 > gpurun_out/prof_layer_halo.log 2>&1; cat gpurun_out/prof_layer_halo.log
ncu --set full --clock-control none --import-source on -k regex:tapgemm -c 4 -o gpurun_out/prof_tg_halo -f \
    python tools/prof_layer.py conv64,conv128 --halo --iters 1 > gpurun_out/ncu_tg_halo.log 2>&1
echo "ncu exit $?"
