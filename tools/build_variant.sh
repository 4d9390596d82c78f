#!/bin/bash
# Developer aid: link a variant of libflucoma_b200.so in which ONE translation unit comes from another source file / extra defines.
# usage: tools/build_variant.sh <out.so> <source replacing csrc/kernels_nmf_tc.cu> [extra nvcc flags...]
#        UNIT=kernels_nmf_tcs tools/build_variant.sh <out.so> <source replacing csrc/$UNIT.cu> [flags...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
UNIT=${UNIT:-kernels_nmf_tc}
OUT=$1; SRC=$2; shift 2
python $ROOT/flucoma-core_b200/build.py > /dev/null
B=$ROOT/flucoma-core_b200/build
TMP=$(mktemp -d)
cp "$SRC" $ROOT/flucoma-core_b200/csrc/.variant_unit.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ \
  --expt-relaxed-constexpr "$@" -c $ROOT/flucoma-core_b200/csrc/.variant_unit.cu -o $TMP/unit.o
rm -f $ROOT/flucoma-core_b200/csrc/.variant_unit.cu
OBJS=$(ls $B/*.o | grep -v "/$UNIT.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "$OUT" $OBJS $TMP/unit.o -lcufft -ldl -Xlinker -rpath,/usr/local/cuda/lib64
rm -rf $TMP
echo "$OUT"
