cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/exp9_pytest.txt
for rep in 1 2; do for o in "tc_chunk32=0" "tc_chunk32=1" "tc_wbuf=3"; do
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e --opt $o 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 $o', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done; done > gpurun_out/exp9_ab.txt 2>&1
python tools/trace_layer.py 1 > gpurun_out/exp9_trace.txt 2>&1
echo done
