#!/bin/bash
# usage: tools/sass_hist.sh <object> <mangled-function-substring>   -> opcode histogram of one kernel
cuobjdump -sass "$1" 2>/dev/null | awk -v f="$2" '/Function :/{on=index($0,f)>0} on' | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn
