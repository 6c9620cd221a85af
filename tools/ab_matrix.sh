#!/bin/bash
# A/B matrix inside ONE gpurun call: every library in $LIBS x every tc_flags value in $FLAGS on a workload.
# usage: LIBS="build/a.so build/b.so" FLAGS="1 49" WL=cfg2s tools/ab_matrix.sh
WL=${WL:-cfg2s}
FLAGS=${FLAGS:-1}
STEPS=${STEPS:-3}
for rep in 1 2; do
  for lib in $LIBS; do
    for f in $FLAGS; do
      WOWSR_LIB=$PWD/$lib python bench.py --workload $WL --steps $STEPS --warmup 2 --no-cpu --opt tc_flags=$f 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'flags=$f', 'ms_per_step', round(d['ms_per_step'],1), 'conv_ms', round(d['roofline']['conv_ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
    done
  done
done
