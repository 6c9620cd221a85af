# Round 2, call 4: rolling kernel after the relaxed hand-back arrive: cycle attribution per role, A/B on cfg2 and cfg5.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/roll_trace.py cfg2s roll_pair=1 > $O/r2c_trace_cfg2s_pair.txt 2>&1
timeout 200 python tools/roll_trace.py cfg2s roll_pair=0 > $O/r2c_trace_cfg2s_single.txt 2>&1
timeout 200 python tools/roll_trace.py cfg5s roll_pair=1 > $O/r2c_trace_cfg5s_pair.txt 2>&1
for v in "roll1:--opt roll=1 --opt roll_pair=0" "roll2:--opt roll=1 --opt roll_pair=1"; do
  name=${v%%:*}; opts=${v#*:}
  timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e $opts > $O/r2c_bench_cfg2_$name.json 2> $O/r2c_bench_cfg2_$name.err
done
timeout 300 python bench.py --workload scene --steps 2 --warmup 1 --no-cpu --no-e2e > $O/r2c_bench_scene_roll2.json 2> $O/r2c_bench_scene_roll2.err
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 --csv \
    --log-file $O/r2c_launches_cfg2s.csv python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2c_ncu_run.log 2>&1
echo done
