cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_assign.py tests/test_gpu_ref_exec.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python tests/fuzz_parity.py --cases 150 --seed 7 2>&1 | tail -3
for flags in "" "--workload crowded" "--workload voc"; do
  timeout 300 python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('[$flags] step %.3f ms  %.0f img/s launches %s' % (d['ms_per_step'], d['value'], d.get('gpu_launches')), {k: round(v['ms'], 4) for k, v in d['kernels'].items()}, d['roofline']['frac'])
"
done
# profiles
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cls_tile_kernel|assign_encode_kernel' --launch-skip 4 --launch-count 2 -o gpurun_out/r02_final_k1_k2 -f python bench.py --no-cpu-baseline --quick --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/ncu_k1k2.log 2>&1
tail -2 gpurun_out/ncu_k1k2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu-baseline --quick --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cls_tile_kernel' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_final_sigma15 -f python bench.py --no-cpu-baseline --quick --steps 3 --warmup 2 --e2e-steps 1 --logit-sigma 1.5 > gpurun_out/ncu_s15.log 2>&1
tail -2 gpurun_out/ncu_s15.log
