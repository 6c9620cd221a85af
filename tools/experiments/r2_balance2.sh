# Round 2: vertical units of conv_last weighted 1.27 — balance at a full cfg5 batch, parity, scene bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/roll_trace.py cfg5b > $O/r3l_trace_cfg5b.txt 2>&1
grep -A6 "conv_last\|conv_hr" $O/r3l_trace_cfg5b.txt | grep "==\|finish time of" 
timeout 900 python -m pytest tests/test_gpu_rrdbnet.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > $O/r3l_bench_scene.json 2> $O/r3l_bench_scene.err
grep -o '"value": [0-9.]*, "unit": "Mpix/s", "n_gpus"' $O/r3l_bench_scene.json; grep -o '"conv_ms_per_step": [0-9.]*' $O/r3l_bench_scene.json; grep -o '"sm_mhz": [0-9.]*' $O/r3l_bench_scene.json
echo done
