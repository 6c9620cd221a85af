# Round 2, call 2: first hardware run of the rolling conv kernel (csrc/roll_kernel.cuh).  One process per step and mode.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out build
O=gpurun_out
[ -x build/mma_2cta_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc tools/mma_2cta_probe.cu -o build/mma_2cta_probe
( timeout 60 build/mma_2cta_probe 96; echo "exit $?"; timeout 60 build/mma_2cta_probe 192; echo "exit $?" ) > $O/r2_2cta_probe.txt 2>&1
for mode in tile roll1 roll2; do
  timeout 300 python tools/roll_check.py layer $mode > $O/r2_roll_layer_$mode.txt 2>&1; echo "exit $?" >> $O/r2_roll_layer_$mode.txt
done
for mode in roll1 roll2; do
  timeout 300 python tools/roll_check.py net $mode > $O/r2_roll_net_$mode.txt 2>&1; echo "exit $?" >> $O/r2_roll_net_$mode.txt
done
for mode in tile roll1 roll2; do
  timeout 200 python tools/roll_check.py perf $mode > $O/r2_roll_perf_$mode.txt 2>&1; echo "exit $?" >> $O/r2_roll_perf_$mode.txt
done
echo done
