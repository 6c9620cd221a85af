cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/exp7_multigpu_check.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/exp7_bench_cfg2_2gpu.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --workload scene > gpurun_out/exp7_bench_scene_2gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/exp7_pytest.txt
echo done
