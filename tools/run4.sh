cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for s in 1.0 1.5; do
  timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 5 --e2e-steps 0 --logit-sigma $s 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('sigma $s: step %.3f ms  %.0f img/s launches %s' % (d['ms_per_step'], d['value'], d.get('gpu_launches')), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['check'])
"
done
