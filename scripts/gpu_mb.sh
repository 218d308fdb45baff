#!/bin/bash
for v in mb6 mb7 mb8; do echo "== $v"; CMGPU_LIB=build/variants/$v.so timeout 300 python scripts/config_bench.py 2>&1 | grep "config 3\|config 4: 4Mi transformed"; done
