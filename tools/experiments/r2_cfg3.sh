# Round 2: launch list of the EDSR x4 workload (cfg3, one 1024x1024 tile) with DRAM bytes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file $O/r2y_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2y_ncu_run.log 2>&1
echo done
