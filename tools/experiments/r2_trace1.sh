# Round 2: phases of a rolling-kernel launch (entry -> barriers/TMEM -> weights -> roles -> exit) on one 128x128 tile and at the cfg5 window size
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/roll_trace.py cfg1 > $O/r3b_trace_cfg1.txt 2>&1
cat $O/r3b_trace_cfg1.txt | grep -v "^   issuer\|^   producer" | head -40
timeout 300 python tools/roll_trace.py cfg5s > $O/r3b_trace_cfg5s.txt 2>&1
grep "phases\|==" $O/r3b_trace_cfg5s.txt | head -30
timeout 600 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | tail -2
echo done
