cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out

python bench.py --workload post4096 --steps 5 --warmup 3 > gpurun_out/exp14_bench_post4096.json 2>/dev/null
echo done
