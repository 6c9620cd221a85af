cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp8_pytest.txt
for rep in 1 2; do for b in 0 1; do
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --opt tc_boustrophedon=$b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 boustrophedon=$b', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done; done > gpurun_out/exp8_ab.txt 2>&1
for b in 0 1; do
  python bench.py --workload scene --steps 1 --warmup 1 --no-cpu --opt tc_boustrophedon=$b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scene boustrophedon=$b', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done >> gpurun_out/exp8_ab.txt 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_net_check.py 1 40 256 > gpurun_out/exp8_memcheck.txt 2>&1
echo done
