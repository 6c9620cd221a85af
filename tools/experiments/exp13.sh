cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_post.py tests/test_io_glue.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/exp13_pytest_post.txt
python tools/post_bench.py > gpurun_out/exp13_post_bench.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/exp13_smoke.txt 2>&1
echo done
