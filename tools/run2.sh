cd $GRAFT_REPO_ROOT
bash tools/sweep.sh base s2 s3 s8t10 w8 w20b1s8 w16b1s5t32 w5b3s6t10 w10b1s8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cls_tile -c 1 -o gpurun_out/r02_tile_focal -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log
