# Round 2, call 9 (2 GPUs): GPU suite with the staged host entry points, sharded scene vs single GPU, cfg5 on 1 and 2 GPUs.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2h_pytest_gpu.txt 2>&1; echo "exit $?" >> $O/r2h_pytest_gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/r2h_multigpu_check_2gpu.txt 2>&1; echo "exit $?" >> $O/r2h_multigpu_check_2gpu.txt
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu > $O/r2h_bench_scene_1gpu.json 2> $O/r2h_bench_scene_1gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2h_bench_scene_2gpu.json 2> $O/r2h_bench_scene_2gpu.err
timeout 300 python bench.py --workload cfg2 --steps 3 --warmup 2 --no-cpu > $O/r2h_bench_cfg2_1gpu.json 2> $O/r2h_bench_cfg2_1gpu.err
echo done
