cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 300 python tools/k1_time.py voc,shapes 0x0,1x1,2x1,1x2,2x2 200
} > gpurun_out/run56.log 2>&1
tail -30 gpurun_out/run56.log
