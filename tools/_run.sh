cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_check.py 2>&1 | grep -E "MGPU_CHECK|Error|error" | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r02b_bench_n8.json 2> gpurun_out/r02b_bench_n8.err
tail -1 gpurun_out/r02b_bench_n8.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=8 value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, 'e2e', d['e2e']['value'], 'strong', d.get('strong_scaling'), 'check', d.get('check'))
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 200 --warmup 10 > gpurun_out/r02b_bench_n4.json 2> gpurun_out/r02b_bench_n4.err
tail -1 gpurun_out/r02b_bench_n4.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=4 value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, 'e2e', d['e2e']['value'], 'strong', d.get('strong_scaling'), 'check', d.get('check'))
"
timeout 300 python bench.py --steps 200 --warmup 10 --quick --no-cpu-baseline --e2e-steps 0 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=1 (same box) value %.0f ms %.4f' % (d['value'], d['ms_per_step']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()})
"
} > gpurun_out/run61.log 2>&1
tail -20 gpurun_out/run61.log
