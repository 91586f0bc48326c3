cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_check.py 2>&1 | grep -E "SANITIZE_RUN_DONE|RACECHECK SUMMARY|Error: Race|and Read|and Write" | head -12
timeout 900 python tests/fuzz_loss.py --cases 1500 --seed 21 2>&1 | tail -2
for flags in "" "--logit-sigma 1.5" "--workload voc"; do
  timeout 300 python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {a: round(v['ms'], 4) for a, v in d['kernels'].items()}, round(d['roofline']['frac'],3))
"
done
