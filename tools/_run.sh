cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for sp in "0 " "0 --pipeline" "8 --pipeline" "16 --pipeline" "24 --pipeline" "32 --pipeline" "40 --pipeline" "16 " "24 "; do
set -- $sp
timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 10 --e2e-steps 0 --quick --spare-sms $1 $2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[spare $sp] step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, round(d['roofline']['frac'],3))
"
done
} > gpurun_out/run49.log 2>&1
tail -50 gpurun_out/run49.log
