#!/bin/bash
# Round-2 queue: energy attribution of the conv kernel.  Every full-size run sits on the 1 kW power cap, so what bounds the
# step is joules, not cycles.  Long enough steps for nvidia-smi to sample power (100 ms period), the
# role-isolation debug flags of the conv kernel (results are garbage, timing and power are real):
#   1 = full kernel, 17 = no TMA loads, 33 = no epilogue, 49 = MMA issue only, 65 = no MMA (TMA + epilogue), 97 = TMA only
# Prints ms per step, SM clock, board power and energy per step (J) for each; the differences attribute the energy to the
# tensor pipe + operand fetch (49), the memory feed (97) and the epilogue (65 - 97).
for f in 1 17 33 49 65 97; do
  python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu --no-e2e --opt tc_flags=$f 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['clocks']
ms = d['ms_per_step']; p = c.get('power_w')
print('tc_flags=%-3d ms_per_step %7.1f  conv_ms %7.1f  sm_mhz %s  power_w %s  energy_J %s  reasons %s' % ($f, ms, d['roofline']['conv_ms_per_step'], c['sm_mhz'], p, round(p * ms / 1e3, 1) if p else None, c['reasons']))"
done
