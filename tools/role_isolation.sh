#!/bin/bash
# Timing-only decomposition of the conv kernel (results are garbage with the debug flags): which role bounds it?
for f in 1 17 33 49 65 81 97 113; do
  echo "tc_flags=$f (bit4 no-TMA, bit5 no-epilogue, bit6 no-MMA)"
  python bench.py --workload cfg2s --steps 2 --warmup 1 --no-cpu --opt tc_flags=$f 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'clk', d['clocks']['sm_mhz'])"
done
