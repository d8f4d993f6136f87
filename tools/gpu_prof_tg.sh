set -u
mkdir -p gpurun_out
python tools/prof_layer.py conv64,conv64plain,conv128,res1x1 --iters 3 > gpurun_out/prof_layer.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tapgemm -c 8 -o gpurun_out/prof_tg python tools/prof_layer.py conv64,conv64plain,conv128,res1x1 --iters 1 > gpurun_out/ncu_tg.log 2>&1
echo "ncu tg exit $?"
CCDM_LA2_SPLIT=0 python tools/prof_linattn.py > gpurun_out/prof_linattn.log 2>&1
CCDM_LA2_SPLIT=1 python tools/prof_linattn.py >> gpurun_out/prof_linattn.log 2>&1
cat gpurun_out/prof_layer.log gpurun_out/prof_linattn.log
