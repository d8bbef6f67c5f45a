#!/bin/bash
LOG2=$1; shift
echo "== default"; python tools/skewbench.py $LOG2 | grep -E "skew +(0|512|1048576) "
for v in "$@"; do echo "== $v"; HPGEOM_B200_LIB=$PWD/variants/lib_$v.so python tools/skewbench.py $LOG2 | grep -E "skew +(0|512|1048576) "; done
