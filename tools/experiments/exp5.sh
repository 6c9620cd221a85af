cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/exp5_dual.py > gpurun_out/exp5_dual.txt 2>&1
echo done
