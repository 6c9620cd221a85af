# Round 2: whole GPU suite + smoke + the bench lines touched by the new post-process kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2z_pytest_gpu.txt 2>&1
tail -4 $O/r2z_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/r2z_smoke.txt 2>&1; tail -2 $O/r2z_smoke.txt
timeout 600 python bench.py --workload post4096 --steps 20 --warmup 5 > $O/r2z_bench_post4096_1gpu.json 2> $O/r2z_bench_post4096_1gpu.err
tail -c 900 $O/r2z_bench_post4096_1gpu.json
timeout 900 python bench.py > $O/r2z_bench_scene_1gpu.json 2> $O/r2z_bench_scene_1gpu.err
tail -c 1500 $O/r2z_bench_scene_1gpu.json
echo done
