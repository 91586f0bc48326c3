cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_assign.py tests/test_gpu_edge.py -x -q 2>&1 | tail -3
timeout 600 python tests/fuzz_parity.py --cases 150 --seed 31 2>&1 | tail -2
V=retinanet_tensorflow_b200/lib/variants
for v in default nocrowd crowd0; do
  if [ $v = default ]; then L=""; else L=$PWD/$V/librndense_$v.so; fi
  echo "== $v"; RNDENSE_LIB=$L timeout 300 python tools/k1_time.py coco,crowded,voc 0x0 100
done
} > gpurun_out/run62.log 2>&1
tail -20 gpurun_out/run62.log
