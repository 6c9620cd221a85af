# Round 2: conv_first with a table of normalised input values; EDSR without the second fp32 copy of the head output
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rrdbnet.py tests/test_gpu_full_size.py tests/test_zz_gpu_edsr_file_entry.py -x -q > $O/r3e_pytest.txt 2>&1
tail -3 $O/r3e_pytest.txt
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu > $O/r3e_bench_cfg3.json 2> $O/r3e_bench_cfg3.err
grep -o '"value": [0-9.]*, "unit": "Mpix/s", "n_gpus"' $O/r3e_bench_cfg3.json; grep -o '"conv_ms_per_step": [0-9.]*' $O/r3e_bench_cfg3.json
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu --no-e2e > $O/r3e_bench_cfg1.json 2> $O/r3e_bench_cfg1.err
grep -o '"conv_ms_per_step": [0-9.]*' $O/r3e_bench_cfg1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_first -c 4 --csv --log-file $O/r3e_first.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r3e_ncu.log 2>&1
grep conv_first $O/r3e_first.csv | tail -2
echo done
