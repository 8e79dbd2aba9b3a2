#!/bin/bash
# Build tuning variants of the library: one .so per line "name|nvcc -D flags" read from stdin (or the file $1),
# into variants/libpmm_<name>.so (git-ignored; travels to the GPU box with gpurun).  Prints ptxas resource usage
# of the lumped kernels for each.
#   scripts/build_variants.sh < variants.txt
cd "$(dirname "$0")/.."
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
build_one() {
    name="${1%%|*}"; flags="${1#*|}"
    $NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v $flags \
        -o variants/libpmm_$name.so laml_b200/csrc/pmm_engine.cu laml_b200/csrc/pmm_ultrametric.cpp > variants/$name.ptxas 2>&1 || { echo "$name: BUILD FAILED"; tail -5 variants/$name.ptxas; return; }
    # resource usage of the lumped kernels
    awk -v n="$name" '/Compiling entry function .*(lumped_outsideILb0|lumped_inside)/ {k=$0; sub(/.*_ZN3pmm[0-9]*/,"",k); sub(/E.*/,"",k)} /Used [0-9]+ registers/ && k {print n": "k": "$0; k=""} /bytes spill/ && k {sp=$0}' variants/$name.ptxas | sed 's/ptxas info    : //'
    grep -B2 -A3 "lumped_outsideILb0" variants/$name.ptxas | grep -E "spill" | head -1 | sed "s/^/$name: /"
}
export -f build_one; export NVCC
grep -v '^#' "${1:-/dev/stdin}" | grep '|' | xargs -P 6 -I{} bash -c 'build_one "$@"' _ {}
