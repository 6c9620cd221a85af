set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/exp1_smi.txt
WOWSR_LIB=$PWD/build/lib_v3.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp1_pytest_v3.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp1_pytest_v0.txt
LIBS="build/lib_base.so build/lib_v0.so build/lib_v1.so build/lib_v2.so build/lib_v3.so" FLAGS="49 1" WL=cfg2s timeout 900 tools/ab_matrix.sh > gpurun_out/exp1_ab.txt 2>&1
for v in base v3; do WOWSR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/trace_layer.py 1 > gpurun_out/exp1_trace_$v.txt 2>&1; done
WOWSR_LIB=$PWD/build/lib_v3.so timeout 200 python tools/trace_layer.py 49 > gpurun_out/exp1_trace_v3_mmaonly.txt 2>&1
echo done
