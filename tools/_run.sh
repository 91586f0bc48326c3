cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fuzz.py tests/test_gpu_ref_exec.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
timeout 600 python tests/fuzz_parity.py --cases 200 --seed 21 2>&1 | tail -3
for sg in 1.0 1.5; do
timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 10 --e2e-steps 0 --quick --logit-sigma $sg 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[coco s$sg] step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['roofline']['frac'], d['check'])
"
done
for wl in crowded voc; do
timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 10 --e2e-steps 0 --quick --workload $wl 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$wl] step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()})
"
done
timeout 300 python bench.py --no-cpu-baseline --steps 60 --warmup 10 --e2e-steps 0 --quick --mode infer 2>&1 | tail -1 | cut -c1-600
} > gpurun_out/run51.log 2>&1
tail -50 gpurun_out/run51.log
