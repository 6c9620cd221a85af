cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/exp10_bench_cfg2_${N}gpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 2 --warmup 2 --workload scene > gpurun_out/exp10_bench_scene_${N}gpu.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/multigpu_check.py > gpurun_out/exp10_multigpu_check_${N}gpu.txt 2>&1
echo done
