# Round 2: shorter minimum segments for small launches — RRDBNet parity tests (cfg1 golden, rolling-grid identity), cfg1 bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rrdbnet.py -x -q > $O/r3c_pytest_rrdbnet.txt 2>&1
tail -3 $O/r3c_pytest_rrdbnet.txt
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 > $O/r3c_bench_cfg1.json 2> $O/r3c_bench_cfg1.err
tail -c 700 $O/r3c_bench_cfg1.json
timeout 300 python tools/roll_trace.py cfg1 > $O/r3c_trace_cfg1.txt 2>&1
grep "phases\|==" $O/r3c_trace_cfg1.txt | head -12
echo done
