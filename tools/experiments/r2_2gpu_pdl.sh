# Round 2 (2 GPUs): sharded scene vs single GPU over NCCL on the end-of-round library (programmatic dependent launch on)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > $O/r4d_multigpu_check_2gpu.txt 2>&1; echo "exit $?" >> $O/r4d_multigpu_check_2gpu.txt
tail -8 $O/r4d_multigpu_check_2gpu.txt
echo done
