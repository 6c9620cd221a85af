# Round 2: EDSR trunk as hi (16-bit copy, updated in place) + lo instead of an fp32 copy — parity, cfg3, and a regression check of rdb.conv5 (same epilogue)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rrdbnet.py tests/test_gpu_full_size.py tests/test_zz_gpu_edsr_file_entry.py tests/test_zz_gpu_geotiff_entry_points.py -x -q > $O/r3h_pytest.txt 2>&1
tail -3 $O/r3h_pytest.txt
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu > $O/r3h_bench_cfg3.json 2> $O/r3h_bench_cfg3.err
grep -o '"value": [0-9.]*, "unit": "Mpix/s", "n_gpus"' $O/r3h_bench_cfg3.json; grep -o '"conv_ms_per_step": [0-9.]*' $O/r3h_bench_cfg3.json
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file $O/r3h_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r3h_ncu_run.log 2>&1
timeout 300 python tools/roll_trace.py cfg2s > $O/r3h_trace_cfg2s.txt 2>&1
grep -A3 "conv5" $O/r3h_trace_cfg2s.txt | grep "==\|issuer" | head -8
echo done
