cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mb in 49152 1800 3500 7000 14000; do
  python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --opt mem_budget_mb=$mb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 budget_mb=$mb', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done > gpurun_out/exp4_budget.txt 2>&1
for mb in 49152 480 960 1920; do
  python bench.py --workload scene --steps 1 --warmup 1 --no-cpu --opt mem_budget_mb=$mb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scene budget_mb=$mb', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done >> gpurun_out/exp4_budget.txt 2>&1
echo done
