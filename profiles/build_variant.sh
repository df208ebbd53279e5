#!/bin/bash
# build_variant.sh <name> [nvcc flags...]: an order-4-only build of libnsem_cuda.so under gpurun_out/variants/<name>/ (with the host library
# next to it) for back-to-back kernel-variant timing: NSEM_LIBDIR=gpurun_out/variants/<name> python bench.py ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
out=$ROOT/variants/$name
mkdir -p "$out"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DNSEM_WITH_NCCL -DNSEM_ONLY_ORDER4_3D "$@" \
    "$ROOT/nebulasem_b200/csrc/nsem_cuda.cu" -o "$out/libnsem_cuda.so" -ldl
cp "$ROOT/nebulasem_b200/lib/libnsem_host.so" "$out/"
