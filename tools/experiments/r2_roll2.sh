# Round 2, call 3: the rolling conv kernel on the full-size workloads (A/B against the tile kernel), the GPU test suite with
# the rolling kernel as the default, and the first ncu passes of the new kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2b_pytest_gpu.txt
for v in "tile:--opt roll=0" "roll1:--opt roll=1 --opt roll_pair=0" "roll2:--opt roll=1 --opt roll_pair=1"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e $opts > $O/r2b_bench_cfg2_$name.json 2> $O/r2b_bench_cfg2_$name.err
done
for v in "tile:--opt roll=0" "roll2:--opt roll=1 --opt roll_pair=1"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 300 python bench.py --workload scene --steps 2 --warmup 1 --no-cpu --no-e2e $opts > $O/r2b_bench_scene_$name.json 2> $O/r2b_bench_scene_$name.err
done
# launch list of one step on 9 windows of 532x532 (serialised / cold cache: compare shares, not absolutes)
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 --csv \
    --log-file $O/r2b_launches_cfg2s.csv python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2b_ncu_run.log 2>&1
# full captures: rdb.conv4 (N = 32, plain epilogue) and rdb.conv5 (N = 64, residual epilogue + identity K-step) of the third RDB
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_roll --launch-skip 13 --launch-count 2 -o $O/r2b_prof_body \
    python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2b_ncu_body.log 2>&1
echo done
