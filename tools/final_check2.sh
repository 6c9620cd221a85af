cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/final2_pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.txt 2>&1
echo done
