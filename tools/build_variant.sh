#!/bin/bash
# Build a variant of the library with extra -D flags for ONE source file (A/B experiments):
#   tools/build_variant.sh NAME FILE.cu "-DFOO=1 -DBAR=2"   ->  tools/bin/libgsm_NAME.so
# The other sources are compiled once into build/obj (rebuilt when stale).
set -eu
NAME=$1; FILE=$2; DEFS=${3:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/build/obj; mkdir -p $OBJ $ROOT/tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC"
unset CC CXX || true
objs=""
for f in $ROOT/gsm_renderer_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b.cu" == "$FILE" ]; then
    o=$OBJ/${b}_$NAME.o
    nvcc $FLAGS $DEFS -c $f -o $o
  else
    o=$OBJ/$b.o
    if [ ! -f $o ] || [ -n "$(find $ROOT/gsm_renderer_b200/csrc $ROOT/include -newer $o \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
      nvcc $FLAGS -c $f -o $o &
    fi
  fi
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/tools/bin/libgsm_$NAME.so $objs
echo built tools/bin/libgsm_$NAME.so
