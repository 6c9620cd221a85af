# What the driver runs at round end, in one call: build check is done on the CPU box; here: GPU tests, smoke, default bench, reference arm.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/final_pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null
echo done
