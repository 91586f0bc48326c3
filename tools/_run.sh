cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
V=retinanet_tensorflow_b200/lib/variants
for v in default r56; do
  if [ $v = default ]; then L=""; else L=$PWD/$V/librndense_$v.so; fi
  echo "== $v"; RNDENSE_LIB=$L timeout 300 python tools/k1_time.py coco,crowded 1x1,2x2,2x1,1x2,3x3,2x3 50
done
timeout 600 python -m pytest tests/test_gpu_assign.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 600 python tests/fuzz_parity.py --cases 150 --seed 11 2>&1 | tail -4
} > gpurun_out/run43.log 2>&1
tail -50 gpurun_out/run43.log
