cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
} > gpurun_out/run48.log 2>&1
tail -20 gpurun_out/run48.log
