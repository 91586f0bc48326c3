cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 300 python tools/k1_time.py coco,crowded 1x1,2x2,2x1,2x3 50
timeout 600 python -m pytest tests/test_gpu_assign.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 600 python tests/fuzz_parity.py --cases 100 --seed 13 2>&1 | tail -4
} > gpurun_out/run45.log 2>&1
tail -50 gpurun_out/run45.log
