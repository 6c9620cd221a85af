#!/bin/bash
# usage: buildvar.sh VAR out.so   (builds conv.cu with -DWOWSR_VAR=VAR, links with the existing ctx/post objects)
set -e
cd /root/repo/sentinel2-super-resolution-poc_b200/csrc
mkdir -p /root/repo/build/var$1
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -gencode arch=compute_100a,code=sm_100a -DWOWSR_VAR=$1 $3 -c conv.cu -o /root/repo/build/var$1/conv.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/build/$2 build/ctx.o build/post.o /root/repo/build/var$1/conv.o -cudart static
