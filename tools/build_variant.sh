#!/bin/bash
# Development aid: build pnfft_b200/lib/variants/<name>.so with extra nvcc defines.  usage: tools/build_variant.sh name -DFOO=1 ...
name="$1"; shift
cd "$(dirname "$0")/../pnfft_b200/csrc"
mkdir -p ../lib/variants
make -j4 OBJDIR=../lib/obj_$name OUT=../lib/variants/$name.so NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v $*" > /tmp/build_$name.log 2>&1 || { tail -20 /tmp/build_$name.log; exit 1; }
grep -A2 "k_gather_zm2IdLb1ELi6ELb1\|k_scatter_zm2IdLb1ELi6ELb0\|k_gather_zm2IdLb1ELi6ELb0" ../lib/obj_$name/api_d.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo
