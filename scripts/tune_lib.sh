#!/bin/bash
# compare alternative builds of libcmgpu.so:  scripts/tune_lib.sh "build/a.so build/b.so" [rays]
for v in $1; do echo "== $v"; CMGPU_LIB=$PWD/$v python scripts/exp_sorted.py ${2:-4194304} 2>&1 | grep -E "input order|lenclass\+morton32:"; done
