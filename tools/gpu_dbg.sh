set -u
mkdir -p gpurun_out
CCDM_SYNC_EACH=1 timeout 300 python tools/prof_forward.py --iters 1 > gpurun_out/dbg_forward.log 2>&1
echo "exit $?" >> gpurun_out/dbg_forward.log
tail -n 30 gpurun_out/dbg_forward.log
