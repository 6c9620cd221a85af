# Round 2: measured-speed balancing of the rolling kernel's work lists — parity (bit-identical across runs), balance of a full cfg5 batch, scene + cfg2 bench A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rrdbnet.py -x -q 2>&1 | tail -2
timeout 600 python tools/roll_trace.py cfg5b > $O/r3m_trace_cfg5b.txt 2>&1
grep "==\|finish time of" $O/r3m_trace_cfg5b.txt | paste - - | awk '{print $2, $3, $4, $(NF-6), $(NF-4), $(NF-2), $NF}'
for a in 1 0; do
  timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu --no-e2e --opt roll_adapt=$a > $O/r3m_bench_cfg2_adapt$a.json 2> $O/r3m_bench_cfg2_adapt$a.err
  echo adapt=$a $(grep -o '"ms_per_step": [0-9.]*' $O/r3m_bench_cfg2_adapt$a.json) $(grep -o '"sm_mhz": [0-9.]*' $O/r3m_bench_cfg2_adapt$a.json) $(grep -o '"power_w": [0-9.]*' $O/r3m_bench_cfg2_adapt$a.json)
done
echo done
