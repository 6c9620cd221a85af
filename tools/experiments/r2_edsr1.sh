# Round 2: EDSR upsampler grouped by sub-pixel phase (plain epilogue with a row pitch) — parity tests, then cfg3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_rrdbnet.py tests/test_gpu_full_size.py tests/test_zz_gpu_edsr_file_entry.py tests/test_zz_gpu_geotiff_entry_points.py -x -q -k "edsr or geotiff" > $O/r3a_pytest_edsr.txt 2>&1
tail -4 $O/r3a_pytest_edsr.txt
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 > $O/r3a_bench_cfg3.json 2> $O/r3a_bench_cfg3.err
tail -c 1200 $O/r3a_bench_cfg3.json
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file $O/r3a_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r3a_ncu_run.log 2>&1
echo done
