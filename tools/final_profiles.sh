# Round-end measurement pass on ONE B200 (run under gpurun): bench lines, ncu launch list with DRAM bytes, ncu --set full captures.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/fp_smi.txt
python bench.py > gpurun_out/fp_bench_cfg2_1gpu.json 2> gpurun_out/fp_bench_cfg2_1gpu.err
python bench.py --workload post4096 --steps 5 --warmup 3 > gpurun_out/fp_bench_post4096.json 2>/dev/null
python bench.py --workload scene --steps 2 --warmup 1 --no-cpu > gpurun_out/fp_bench_scene_1gpu.json 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/fp_bench_reference.json 2>/dev/null
python tools/post_bench.py > gpurun_out/fp_post_bench.txt 2>&1
python tools/trace_layer.py 1 > gpurun_out/fp_trace_layers.txt 2>&1
# launch list (first step only: -c 360), serialised / cold cache: compare shares, not absolutes
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 --csv \
    --log-file gpurun_out/fp_launches_cfg2s.csv python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/fp_ncu_run.log 2>&1
# full captures: rdb.conv4 (N=32, plain epilogue), rdb.conv5 (N=64, residual epilogue + identity K-step), conv_hr, post kernels
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc --launch-skip 13 --launch-count 2 -o gpurun_out/fp_prof_body \
    python bench.py --workload cfg2s --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/fp_ncu_body.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"post_apply|clahe_hist" --launch-count 2 -o gpurun_out/fp_prof_post \
    python bench.py --workload post4096 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/fp_ncu_post.log 2>&1
echo done
