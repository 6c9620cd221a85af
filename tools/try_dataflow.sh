cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 45 python tools/try_dataflow_perf.py > gpurun_out/try_dataflow_perf.txt 2>&1
echo "exit $?" >> gpurun_out/try_dataflow_perf.txt
echo done
