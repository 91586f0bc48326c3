cd $GRAFT_REPO_ROOT
V=$PWD/retinanet_tensorflow_b200/lib/variants
for lib in "" $V/librndense_ilp2.so $V/librndense_ilp3.so $V/librndense_ilp5.so; do
echo "=== lib: $lib"
for flags in "" "--logit-sigma 1.5" "--logits-dtype bf16" "--cls-mode bce_dice"; do
  RNDENSE_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --quick --steps 60 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, round(d['roofline']['frac'],3))
"
done
done
