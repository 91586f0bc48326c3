cd $GRAFT_REPO_ROOT
V=$PWD/retinanet_tensorflow_b200/lib/variants
for lib in "" $V/librndense_w7.so $V/librndense_w7r80.so; do
echo "=== lib: $lib"
RNDENSE_LIB=$lib python tools/overlap_probe.py coco 2>&1 | head -12
for flags in "" "--no-pipeline" "--workload crowded"; do
  RNDENSE_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s launches %s' % (d['ms_per_step'], d['value'], d.get('gpu_launches')), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['roofline']['frac'])
"
done
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
