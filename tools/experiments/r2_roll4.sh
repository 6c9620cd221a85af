# Round 2, call 5: rolling kernel with schedule-independent ring slots (row v -> slot v % RING) and the residual prefetch.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for mode in roll1 roll2; do
  timeout 300 python tools/roll_check.py layer $mode > $O/r2d_roll_layer_$mode.txt 2>&1; echo "exit $?" >> $O/r2d_roll_layer_$mode.txt
  timeout 300 python tools/roll_check.py net $mode > $O/r2d_roll_net_$mode.txt 2>&1; echo "exit $?" >> $O/r2d_roll_net_$mode.txt
done
timeout 600 python -m pytest tests -m gpu -q > $O/r2d_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2d_pytest_gpu.txt
timeout 200 python tools/roll_trace.py cfg2s roll_pair=1 > $O/r2d_trace_cfg2s_pair.txt 2>&1
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2d_bench_cfg2_roll2.json 2> $O/r2d_bench_cfg2_roll2.err
timeout 300 python bench.py --workload scene --steps 2 --warmup 1 --no-cpu --no-e2e > $O/r2d_bench_scene_roll2.json 2> $O/r2d_bench_scene_roll2.err
echo done
