#!/bin/bash
# scratch/sassloop.sh <lib.so> <mangled-name-prefix>: opcode histogram of a kernel's SASS (static), FP64 share
cuobjdump -sass $1 | awk '/Function : /{f=$3} {print f "\t" $0}' | grep "^$2" | cut -f2- | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#^\s+/\*[0-9a-f]+\*/\s+##; s#;.*##' > /tmp/k.sass
wc -l < /tmp/k.sass
awk '{op=$1; if (op ~ /^@/) op=$2; split(op,a,"."); print a[1]}' /tmp/k.sass | sort | uniq -c | sort -rn | head -${3:-16}
