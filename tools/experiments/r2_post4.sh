# Round 2, post-process: LUT values converted on the ALU pipe (I2FP) instead of the XU pipe (I2F.U16)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python bench.py --workload post4096 --steps 30 --warmup 5 --no-cpu --no-e2e 2>> $O/r3f.err | grep -o '"ms_per_step": [0-9.]*'; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"post_march" -c 2 --csv --log-file $O/r3f_ncu_post_metrics.csv \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r3f_ncu_run.log 2>&1
grep post_march $O/r3f_ncu_post_metrics.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -4
echo done
