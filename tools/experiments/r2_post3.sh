# Round 2, post-process: several strips per CTA (shared tables) — bit-exactness, then a sweep of (threads per strip, strips per CTA),
# and the histogram kernel with plain shared atomics.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q > $O/r2w_pytest_post.txt 2>&1
tail -3 $O/r2w_pytest_post.txt
run() { timeout 300 python bench.py --workload post4096 --steps 20 --warmup 5 --no-cpu --no-e2e "$@" 2>> $O/r2w_bench.err | grep -o '"ms_per_step": [0-9.]*'; }
echo default $(run)
for cfg in "64 8" "64 4" "96 5" "128 4" "128 3" "128 2" "160 3" "256 2" "256 1"; do set -- $cfg; echo nt=$1 groups=$2 $(run --opt post_nt=$1 --opt post_groups=$2); done
echo hist_match=2 $(run --opt hist_match=2)
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"post_march|clahe_hist" -c 4 --csv --log-file $O/r2w_ncu_post_metrics.csv \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e --opt hist_match=2 > $O/r2w_ncu_run.log 2>&1
echo done
