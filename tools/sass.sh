#!/bin/bash
# usage: tools_sass.sh <lib.so> <substring of kernel name>   -> /tmp/k.sass + histogram
LIB=$1; PAT=$2
cuobjdump -sass $LIB | awk -v pat="$PAT" '/Function :/ {p = index($0, pat) > 0} p {print}' > /tmp/k.sass
grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-25}
