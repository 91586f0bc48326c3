#!/bin/bash
# sweep.sh VARIANT... : bench the headline workload with each tuning build (see build_variant.sh)
cd $GRAFT_REPO_ROOT
V=retinanet_tensorflow_b200/lib/variants
for v in "$@"; do
  for s in 1.0 1.5; do
    RNDENSE_LIB=$PWD/$V/librndense_$v.so timeout 300 python bench.py --no-cpu-baseline --steps 40 --warmup 5 --e2e-steps 0 --logit-sigma $s 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('$v sigma $s: step %.3f ms  cls_loss %.4f ms frac %.3f' % (d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']))
except Exception as e:
    print('$v sigma $s: FAILED', e)
"
  done
done
