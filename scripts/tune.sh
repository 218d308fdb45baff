#!/bin/bash
# run the sorted/unsorted experiment for several values of an env knob:  scripts/tune.sh CMGPU_GANG "1 4 8 12 16 24"
for v in $2; do echo "== $1=$v"; env $1=$v python scripts/exp_sorted.py ${3:-4194304} 2>&1 | grep -E "input order|lenclass\+morton32:"; done
