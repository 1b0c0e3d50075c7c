#!/bin/bash
# usage: sass_ops.sh <cubin|so> <kernel-name-substring>  -> compact list of float/int-convert ops of that kernel
f=$1; k=$2
name=$(cuobjdump -sass "$f" | grep "Function :" | grep "$k" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$name" "$f" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\*[0-9a-f]{4}\*/##; s#/\* 0x[0-9a-f]+ \*/##; s/\s+/ /g' | grep -E "FFMA|FMUL|FADD|MUFU|F2I|I2F|FRND|DFMA|DADD|DMUL|F2F|FMNMX|FSETP|CALL|BRA|FSEL|FCHK"
