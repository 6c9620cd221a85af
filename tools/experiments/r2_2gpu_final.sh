# Round 2 (2 GPUs): sharded scene vs single GPU over NCCL with the strip-march post-process, cfg5 on 2 GPUs, reference arm under torchrun
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > $O/r3d_multigpu_check_2gpu.txt 2>&1; echo "exit $?" >> $O/r3d_multigpu_check_2gpu.txt
tail -6 $O/r3d_multigpu_check_2gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r3d_bench_scene_2gpu.json 2> $O/r3d_bench_scene_2gpu.err
tail -c 1300 $O/r3d_bench_scene_2gpu.json
echo done
