#!/bin/bash
# Development aid: build the CUDA library with extra nvcc flags into swraster-viewer_b200/lib_<name>/ (next to a copy of the host
# library), so that several variants can be A/B-timed in ONE gpurun call: SWR_LIB_DIR=swraster-viewer_b200/lib_<name> python ...
# usage: tools/build_variant.sh <name> [extra nvcc flags...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/swraster-viewer_b200/lib_$name
mkdir -p "$out"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
    -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ "$@" -o "$out/libswr_b200.so" "$root/swraster-viewer_b200/csrc/swr_api.cu"
/usr/bin/g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -fopenmp -o "$out/libswr_host.so" "$root/swraster-viewer_b200/host/swr_host_c.cpp" \
    -L"$out" -lswr_b200 -lz -ldl -Wl,-rpath,'$ORIGIN'
echo "built $out"
