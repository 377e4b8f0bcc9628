#!/bin/bash
# Dev tool: build a variant of libwsocean.so with different CTA tilings.  usage: tune_build.sh NAME -DWSO_TUNE_CP10=2 ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-O2 --shared -cudart static -ccbin /usr/bin/g++ \
  -I watersurfacerendering_b200/csrc -I include "$@" -o build/variants/libwsocean_$name.so \
  watersurfacerendering_b200/csrc/wso_kernels.cu watersurfacerendering_b200/csrc/wso_kernels2.cu watersurfacerendering_b200/csrc/wso_slab_kernels.cu watersurfacerendering_b200/csrc/wso_prepare_kernels.cu watersurfacerendering_b200/csrc/wso_api.cu watersurfacerendering_b200/csrc/wso_slab.cu watersurfacerendering_b200/csrc/wso_host_prepare.cpp
echo built build/variants/libwsocean_$name.so
