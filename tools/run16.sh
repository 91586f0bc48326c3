cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','stats_exchange','assign_one_step_ahead')})
print('kernels', {k: round(v['ms'],4) for k,v in d['kernels'].items()})
print('strong', {k:v for k,v in (d.get('strong_scaling') or {}).items() if k!='note'})
print('check', d['check'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
print('sigma15', {k:d['roofline_sigma15'][k] for k in ('frac','ms_per_launch','ms_per_step')})
PY
done
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --api none 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=1 step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {a: round(v['ms'], 4) for a, v in d['kernels'].items()}, d['roofline']['frac'])
"
