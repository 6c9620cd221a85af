# Round 2, call 12: final single-GPU evidence — GPU suite + smoke, bench lines, ncu --set full of one RDB at the FULL cfg2 size.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2k_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2k_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.txt 2>&1; echo "exit $?" >> $O/r2k_smoke.txt
timeout 600 python bench.py > $O/r2k_bench_scene_1gpu.json 2> $O/r2k_bench_scene_1gpu.err
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 > $O/r2k_bench_cfg2_1gpu.json 2> $O/r2k_bench_cfg2_1gpu.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2k_bench_cfg3_1gpu.json 2> $O/r2k_bench_cfg3_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_roll --launch-skip 10 --launch-count 5 -o $O/r2k_prof_rdb_cfg2 \
    python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2k_ncu_rdb.log 2>&1
echo done
