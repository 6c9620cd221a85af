# Round-2 queue — the experiments prepared on the CPU box after round 1's GPU budget was spent (none of them has run on
# hardware yet).  ONE gpurun call, every step under its own timeout so a hang costs seconds, not the box:
#   gpurun --timeout 900 -- 'bash tools/experiments/r2_queue.sh'
# 1. tcgen05 collector / weight-stationary microbenchmark (tools/mma_ws_bench.cu): one (mode, N) per process, because an
#    illegal shape faults the context.  Question: can B (the weights) be held across the R + 2 input rows of a tile?
# 2. fused conv4+conv5 launch (csrc/sched_kernel.cuh): correctness on small shapes, then trunk time on 25 windows of 276x276
#    against the layer-by-layer path, then CTA 0's per-task trace.
# 3. DRAM traffic of both paths (is the dense buffer really found in L2?): ncu dram bytes over one forward of 2 blocks.
# 4. the new HSV vegetation-mask kernel's tests (they also run in the round-end pytest -m gpu).
# 5. CTA-pair MMA rate (tools/mma_2cta_bench.cu): does cta_group::2 reach max(N/2, 32 + N/8) cycles (48 at N_eff = 96)?
# 6. TMA zero-stride probe (tools/tma_stride0_probe.cu): can the nearest-x2 upsample be folded into the consumer's tensor map?
#    and, if so, the kernel built on it (csrc/ups_kernel.cuh, option tail_fold_upsample): bit-identical output, tail time.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out build
O=gpurun_out
[ -x build/mma_ws_bench ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc tools/mma_ws_bench.cu -o build/mma_ws_bench
[ -x build/mma_2cta_bench ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc tools/mma_2cta_bench.cu -o build/mma_2cta_bench
[ -x build/tma_stride0_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc tools/tma_stride0_probe.cu -o build/tma_stride0_probe
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
( for mode in 0 1 2 3; do for n in 32 64 96 128 192; do timeout 60 build/mma_ws_bench $mode $n || echo "mode=$mode N=$n exit $?"; done; done ) > $O/r2_mma_ws_bench.txt 2>&1
( for n in 32 64 96 128 192; do timeout 60 build/mma_2cta_bench $n || echo "N=$n exit $?"; done ) > $O/r2_mma_2cta_bench.txt 2>&1
timeout 30 build/tma_stride0_probe > $O/r2_tma_stride0_probe.txt 2>&1; echo "exit $?" >> $O/r2_tma_stride0_probe.txt
# (the steps that drove the fused conv4+conv5 launch, the dataflow trunk and the opt-in pytest forms were removed with those kernels
#  after this queue had measured them: profiles/r02_queue_fused_*.txt, r02_queue_dram_compare.txt, r02_queue_fold_upsample.txt)
timeout 120 python -m pytest tests/test_zz_gpu_green_mask.py -x -q -m gpu > $O/r2_green_mask_pytest.txt 2>&1
echo done
