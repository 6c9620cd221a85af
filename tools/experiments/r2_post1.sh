# Round 2, post-process call 3 (after the pipelined exchange): strip-march pass-B kernel — bit-exactness (both kernels, every shape), then the 4096x4096 workload
# on both kernels and instruction counts from ncu.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q > $O/r2v_pytest_post.txt 2>&1
tail -5 $O/r2v_pytest_post.txt
for pk in 1 0; do
  timeout 300 python bench.py --workload post4096 --steps 20 --warmup 5 --no-cpu --opt post_kernel=$pk > $O/r2v_bench_post4096_pk$pk.json 2> $O/r2v_bench_post4096_pk$pk.err
  tail -c 600 $O/r2v_bench_post4096_pk$pk.json
done
for nt in 128; do
  timeout 300 python bench.py --workload post4096 --steps 20 --warmup 5 --no-cpu --no-e2e --opt post_nt=$nt > $O/r2v_bench_post4096_nt$nt.json 2>> $O/r2v_bench_nt.err
done
grep -h -o '"ms_per_step": [0-9.]*' $O/r2v_bench_post4096_nt*.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"post_march|post_apply|clahe_hist" -c 8 --csv --log-file $O/r2v_ncu_post_metrics.csv \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2v_ncu_run.log 2>&1
echo done
