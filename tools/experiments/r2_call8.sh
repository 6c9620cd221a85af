# Round 2, call 8: full GPU suite, folded-upsample A/B under the power cap, the bench lines of every workload.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/r2g_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2g_pytest_gpu.txt
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e --opt roll_ups=0 > $O/r2g_bench_cfg2_tileups.json 2> $O/r2g_bench_cfg2_tileups.err
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2g_bench_cfg2.json 2> $O/r2g_bench_cfg2.err
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e --opt roll_ups=0 > $O/r2g_bench_cfg2_tileups2.json 2> $O/r2g_bench_cfg2_tileups2.err
( time timeout 600 python bench.py ) > $O/r2g_bench_default.json 2> $O/r2g_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r2g_bench_reference.json 2> $O/r2g_bench_reference.err
timeout 200 python bench.py --workload cfg1 --steps 20 --warmup 5 > $O/r2g_bench_cfg1.json 2> $O/r2g_bench_cfg1.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2g_bench_cfg3.json 2> $O/r2g_bench_cfg3.err
timeout 200 python bench.py --workload post4096 --steps 10 --warmup 3 > $O/r2g_bench_post4096.json 2> $O/r2g_bench_post4096.err
echo done
