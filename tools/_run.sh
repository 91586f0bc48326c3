cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_ref_exec.py tests/test_gpu_decode.py tests/test_gpu_golden.py tests/test_shapes.py -x -q 2>&1 | tail -5
timeout 600 python tools/prof_modules.py 2>&1 | grep "per image"
} > gpurun_out/run53.log 2>&1
tail -20 gpurun_out/run53.log
