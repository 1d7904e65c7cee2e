#!/bin/bash
# Builds the (not shipped) split-band attention experiment into build_variants/lib_<name>.so without touching the tree:
#   tools/build_split_variant.sh <name> [nvcc -D flags...]
# The experiment source is tools/experiments/attn_tcgen05_split.cu (a variant of csrc/attn_tcgen05.cu, never built by the Makefile).
set -e
name=$1; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
csrc=$root/cmt-cooperative-perception_b200/csrc
mkdir -p $root/build_variants /tmp/split_exp
cp $root/tools/experiments/attn_tcgen05_split.cu /tmp/split_exp/attn_split.cu
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -I$csrc "$@" \
     -c /tmp/split_exp/attn_split.cu -o /tmp/split_exp/${name}.o 2> $root/build_variants/${name}.ptxas.log
others=$(ls $csrc/*.o | grep -v "attn_tcgen05.o$")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/build_variants/lib_${name}.so /tmp/split_exp/${name}.o $others -cudart static
grep -E "spill" $root/build_variants/${name}.ptxas.log | sort | uniq -c | head -4
