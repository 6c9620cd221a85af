set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp3_pytest_s8.txt
WOWSR_LIB=$PWD/build/lib_s16.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp3_pytest_s16.txt
LIBS="build/lib_v0.so build/lib_e16.so build/lib_s8.so build/lib_s16.so" FLAGS="1" WL=cfg2s timeout 600 tools/ab_matrix.sh > gpurun_out/exp3_ab.txt 2>&1
for v in s8i s16i; do WOWSR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/trace_layer.py 1 > gpurun_out/exp3_trace_$v.txt 2>&1; done
echo done
