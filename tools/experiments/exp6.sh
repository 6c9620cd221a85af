cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/exp6_pytest.txt
python tools/gpu_net_check.py 2 48 256 > gpurun_out/exp6_netcheck.txt 2>&1
for rep in 1 2; do for hl in 0 1; do
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --opt trunk_hilo=$hl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 trunk_hilo=$hl', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done; done > gpurun_out/exp6_ab.txt 2>&1
echo done
