cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','stats_exchange')})
print('kernels', {k: round(v['ms'],4) for k,v in d['kernels'].items()})
print('strong', d.get('strong_scaling'))
print('check', d['check'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 100 --warmup 5 --quick --pipeline 2>&1 | tail -1 | cut -c1-400
