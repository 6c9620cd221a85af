# Round 2, call 11 (8 GPUs): sharded scene vs single GPU over NCCL, cfg5 strong scaling and cfg2 weak scaling at 8.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > $O/r2j_multigpu_check_8gpu.txt 2>&1; echo "exit $?" >> $O/r2j_multigpu_check_8gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2j_bench_scene_8gpu.json 2> $O/r2j_bench_scene_8gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 2 > $O/r2j_bench_scene_4gpu.json 2> $O/r2j_bench_scene_4gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --workload cfg2 --steps 5 --warmup 3 > $O/r2j_bench_cfg2_8gpu.json 2> $O/r2j_bench_cfg2_8gpu.err
echo done
