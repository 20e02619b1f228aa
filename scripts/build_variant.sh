#!/bin/bash
# build_variant.sh <tag> <extra nvcc flags...>: liboov_b200.so with tc_score.cu compiled with the extra flags -> build/variants/lib_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
python -c "import __graft_entry__ as g; g.build()" > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c improving-inductive-oov-recsys_b200/csrc/tc_score.cu -o build/variants/tc_score_$tag.o
objs=$(ls improving-inductive-oov-recsys_b200/build/*.o | grep -v tc_score.o)
nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -o build/variants/lib_$tag.so $objs build/variants/tc_score_$tag.o -lcuda
echo built build/variants/lib_$tag.so
