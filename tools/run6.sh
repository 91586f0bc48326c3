cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for flags in "" "--no-pipeline" "--logit-sigma 1.5" "--workload crowded" "--workload crowded --no-pipeline" "--workload voc" "--cls-mode bce_dice" "--logits-dtype bf16"; do
  timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 5 --e2e-steps 0 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s launches %s' % (d['ms_per_step'], d['value'], d.get('gpu_launches')), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['roofline']['frac'], d['check'])
"
done
timeout 900 python tests/fuzz_loss.py --cases 500 --seed 5 2>&1 | tail -4
timeout 600 python tests/fuzz_select.py --cases 200 --seed 5 2>&1 | tail -3
