#!/bin/bash
# usage: tools/variant.sh <name> <file.cu> [-DFLAG ...]   -> variants/lib_<name>.so = libhpgb.so with ONE
# translation unit rebuilt with extra flags (A/B timing of kernel variants via HPGEOM_B200_LIB)
set -e
NAME=$1; SRC=$2; shift 2
cd "$(dirname "$0")/../hpgeom_b200/csrc"
mkdir -p ../../variants /tmp/vb_$NAME
nvcc -diag-suppress 177 -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c $SRC.cu -o /tmp/vb_$NAME/$SRC.o
OBJS=""
for o in build/*.o; do b=$(basename $o .o); if [ "$b" = "$SRC" ]; then OBJS="$OBJS /tmp/vb_$NAME/$SRC.o"; else OBJS="$OBJS $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared $OBJS -o ../../variants/lib_$NAME.so -lcudart
echo built variants/lib_$NAME.so
