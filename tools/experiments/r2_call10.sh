# Round 2, call 10: early accumulator hand-back in the rolling epilogue — correctness, cycle attribution, same-box A/B.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for mode in roll1 roll2; do
  timeout 300 python tools/roll_check.py layer $mode > $O/r2i_roll_layer_$mode.txt 2>&1; echo "exit $?" >> $O/r2i_roll_layer_$mode.txt
done
timeout 300 python tools/roll_check.py net roll2 > $O/r2i_roll_net_roll2.txt 2>&1; echo "exit $?" >> $O/r2i_roll_net_roll2.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2i_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2i_pytest_gpu.txt
timeout 200 python tools/roll_trace.py cfg2s roll_pair=1 > $O/r2i_trace_cfg2s_pair.txt 2>&1
WOWSR_LIB=$PWD/build/libwowsr_base.so timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2i_bench_cfg2_base.json 2> $O/r2i_bench_cfg2_base.err
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2i_bench_cfg2_new.json 2> $O/r2i_bench_cfg2_new.err
WOWSR_LIB=$PWD/build/libwowsr_base.so timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2i_bench_cfg2_base2.json 2> $O/r2i_bench_cfg2_base2.err
timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu --no-e2e > $O/r2i_bench_cfg2_new2.json 2> $O/r2i_bench_cfg2_new2.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu > $O/r2i_bench_cfg3.json 2> $O/r2i_bench_cfg3.err
echo done
