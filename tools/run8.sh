cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
V=$PWD/retinanet_tensorflow_b200/lib/variants
for flags in "" "--no-pipeline" "--workload crowded" "--logits-dtype bf16"; do
  timeout 300 python bench.py --no-cpu-baseline --quick --steps 60 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s launches %s' % (d['ms_per_step'], d['value'], d.get('gpu_launches')), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['roofline']['frac'], d['e2e'])
"
done
RNDENSE_LIB=$V/librndense_m2.so timeout 300 python bench.py --no-cpu-baseline --quick --steps 60 --warmup 5 --e2e-steps 3 --logits-dtype bf16 2>&1 | tail -1 | cut -c1-300
timeout 900 python tests/fuzz_loss.py --cases 500 --seed 5 2>&1 | grep -v "cases ok" | tail -12
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
