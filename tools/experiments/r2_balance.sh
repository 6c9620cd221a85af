# Round 2: load balance of the rolling kernel's work list at a full cfg5 batch (144 windows of 276x276)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/roll_trace.py cfg5b > $O/r3k_trace_cfg5b.txt 2>&1
grep "==\|finish\|LR-res" $O/r3k_trace_cfg5b.txt | head -40
echo done
