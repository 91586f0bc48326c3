set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_gpu_decode.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15
for s in 1.0 1.5 2.0; do
for t in 1 0; do
  echo "== sigma $s tile $t"
  timeout 300 python bench.py --no-cpu-baseline --steps 50 --warmup 5 --e2e-steps 0 --logit-sigma $s --opt tile_kernel=$t 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline'], d.get('kernels'))
"
done
done
