# Round 2, call 15: ncu launch list of one step at the DEFAULT workload's window size (25 windows of 276x276), full capture of the HR tail.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 --csv \
    --log-file $O/r2n_launches_cfg5s.csv python bench.py --workload cfg5s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2n_ncu_run.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_roll|post_apply|clahe_hist" --launch-skip 345 --launch-count 7 -o $O/r2n_prof_tail \
    python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2n_ncu_tail.log 2>&1
echo done
