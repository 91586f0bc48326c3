cd $GRAFT_REPO_ROOT
for flags in "--workload voc" "--workload voc --graph" "--workload crowded" "--workload shapes --batch 1 --flip" "--workload shapes --batch 1 --flip --graph" "--cls-mode bce_dice" "--logits-dtype bf16" "--logits-dtype bf16 --cls-mode bce_dice" "--mode infer" "--logit-sigma 2.0"; do
  timeout 300 python bench.py --no-cpu-baseline --quick --steps 100 --warmup 5 --e2e-steps 3 $flags 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d.get('kernels') or {}
print('[$flags] step %.4f ms  %.0f img/s' % (d['ms_per_step'], d['value']), {a: round(v['ms'], 4) for a, v in k.items()}, round((d.get('roofline') or {}).get('frac') or 0, 3), d.get('gpu_launches'))
"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cls_tile_kernel' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_bf16 -f python bench.py --no-cpu-baseline --quick --steps 3 --warmup 2 --e2e-steps 1 --logits-dtype bf16 > gpurun_out/ncu_bf16.log 2>&1
tail -2 gpurun_out/ncu_bf16.log
