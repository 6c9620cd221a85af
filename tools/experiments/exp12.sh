cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_io_glue.py tests/test_gpu_rrdbnet.py::test_kernel_option_paths_agree -m gpu -x -q -s 2>&1 | tail -15 ) > gpurun_out/exp12_pytest_full.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/exp12_smoke.txt 2>&1
echo done
