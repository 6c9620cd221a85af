set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 1 65 129 17; do WOWSR_LIB=$PWD/build/lib_v4.so timeout 200 python tools/trace_layer.py $f > gpurun_out/exp2_trace_v4_f$f.txt 2>&1; done
WOWSR_LIB=$PWD/build/lib_v12.so timeout 200 python tools/trace_layer.py 1 > gpurun_out/exp2_trace_v12_f1.txt 2>&1
LIBS="build/lib_v0.so build/lib_v8.so" FLAGS="1" WL=cfg2s timeout 600 tools/ab_matrix.sh > gpurun_out/exp2_ab.txt 2>&1
WOWSR_LIB=$PWD/build/lib_v8.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp2_pytest_v8.txt
echo done
