# Round 2: is cfg1 (one 128x128 tile, 354 launches in 6.8 ms) bound by the host feeding launches or by per-kernel latency on the device?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu > $O/r2x_bench_cfg1.json 2> $O/r2x_bench_cfg1.err
tail -c 400 $O/r2x_bench_cfg1.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2x_launches_cfg1.csv \
    python bench.py --workload cfg1 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2x_ncu_run.log 2>&1
echo done
