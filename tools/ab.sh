#!/bin/bash
# usage: tools/ab.sh <log2> <which> [variant names...]   (run on the GPU box) times the default lib and each variant
LOG2=$1; WHICH=$2; shift 2
echo "== default"; python tools/quickbench.py $LOG2 $WHICH
for v in "$@"; do echo "== $v"; HPGEOM_B200_LIB=$PWD/variants/lib_$v.so python tools/quickbench.py $LOG2 $WHICH; done
