cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -2 gpurun_out/r02_bench_ref.err
tail -c 600 gpurun_out/r02_bench_ref.json
