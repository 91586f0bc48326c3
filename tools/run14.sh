cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error" | tail -60
timeout 600 python tests/fuzz_loss.py --cases 300 --seed 9 2>&1 | grep -v "cases ok" | tail -4
for flags in "" "--logit-sigma 1.5"; do
  timeout 300 python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, round(d['roofline']['frac'],3))
"
done
