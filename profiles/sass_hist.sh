#!/bin/bash
# usage: sass_hist.sh <object> <mangled-or-substring of kernel name>: SASS opcode histogram
obj=$1; pat=$2
fn=$(cuobjdump -sass $obj | grep -E "Function : " | grep -- "$pat" | head -1 | awk '{print $3}')
echo "function: $fn" | c++filt
cuobjdump -sass -fun "$fn" $obj | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -${3:-25}
echo total: $(cuobjdump -sass -fun "$fn" $obj | grep -cE "^\s+/\*[0-9a-f]{4}\*/")
