# Round 2, call 6: folded upsample inside the rolling kernel; ncu of the epilogue-bound layers; new parity tests.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/roll_check.py net roll2 > $O/r2e_roll_net_roll2.txt 2>&1; echo "exit $?" >> $O/r2e_roll_net_roll2.txt
timeout 300 python tools/roll_check.py net roll1 > $O/r2e_roll_net_roll1.txt 2>&1; echo "exit $?" >> $O/r2e_roll_net_roll1.txt
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $O/r2e_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2e_pytest_gpu.txt
timeout 200 python tools/roll_trace.py cfg2s roll_pair=1 > $O/r2e_trace_cfg2s_pair.txt 2>&1
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2e_bench_cfg2.json 2> $O/r2e_bench_cfg2.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_roll --launch-skip 8 --launch-count 7 -o $O/r2e_prof_body \
    python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2e_ncu_body.log 2>&1
echo done
