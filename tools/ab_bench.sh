#!/bin/bash
# A/B of two libwowsr builds inside ONE gpurun call (clocks differ between boxes, so compare only within a call).
# usage: tools/ab_bench.sh build/lib_old.so build/lib_new.so [workload]
WL=${3:-cfg2}
for rep in 1 2; do
  for lib in "$1" "$2"; do
    WOWSR_LIB=$PWD/$lib python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'ms_per_step', round(d['ms_per_step'],1), 'conv TF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
  done
done
