#!/bin/bash
# Builds kernel variants HERE (nvcc cross-compiles without a GPU) as csrc/variants/libtess_<name>.so so that one short
# gpurun call can time them all:  tools/build_variants.sh name1="<flags>" name2="<flags>" ...
# then on the GPU box:            tools/run_variants.sh [bench args]
set -e
cd "$(dirname "$0")/../vk_tessellated_clusters_b200/csrc"
mkdir -p variants
rm -f variants/*.so
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  (
    tmp=$(mktemp -d)
    nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -c tc_kernels.cu -o $tmp/k.o
    nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -c tc_api.cu -o $tmp/a.o
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libtess_$name.so $tmp/k.o $tmp/a.o tc_clusterize.o
    rm -rf $tmp
    echo "built $name ($flags)"
  ) &
done
wait
