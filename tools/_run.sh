cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 1200 python tests/fuzz_tail.py --cases 800 --seed 5 2>&1 | tail -12
} > gpurun_out/run58.log 2>&1
tail -20 gpurun_out/run58.log
