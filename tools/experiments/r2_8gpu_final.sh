# Round 2 (8 GPUs): sharded scene vs single GPU over NCCL and the cfg5 strong-scaling line on the end-of-round kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > $O/r3x_multigpu_check_8gpu.txt 2>&1; echo "exit $?" >> $O/r3x_multigpu_check_8gpu.txt
tail -3 $O/r3x_multigpu_check_8gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > $O/r3x_bench_scene_8gpu.json 2> $O/r3x_bench_scene_8gpu.err
python - <<PY
import json
j=json.loads(open("$O/r3x_bench_scene_8gpu.json").read().strip().split("\n")[-1])
print("scene 8 GPUs", round(j["value"],1), "ms", round(j["ms_per_step"],1), "e2e", round(j["e2e"]["value"],1), "roofline", round(j["roofline"]["frac"],4), j["clocks"])
PY
echo done
