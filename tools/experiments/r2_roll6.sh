# Round 2, call 7: accumulator-pair wait moved behind the group's commit; full GPU suite; tile-kernel yardstick in the same call.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/r2f_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2f_pytest_gpu.txt
timeout 200 python tools/roll_trace.py cfg2s roll_pair=1 > $O/r2f_trace_cfg2s_pair.txt 2>&1
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e --opt roll=0 > $O/r2f_bench_cfg2_tile.json 2> $O/r2f_bench_cfg2_tile.err
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2f_bench_cfg2.json 2> $O/r2f_bench_cfg2.err
timeout 300 python bench.py --workload scene --steps 2 --warmup 1 --no-cpu --no-e2e > $O/r2f_bench_scene.json 2> $O/r2f_bench_scene.err
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 --csv \
    --log-file $O/r2f_launches_cfg2s.csv python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2f_ncu_run.log 2>&1
echo done
