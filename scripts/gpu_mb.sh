#!/bin/bash
for sm in 0 14000 18000 28000 37000 56000 75000 113000; do echo "== smem $sm"; CMGPU_AAS_SMEM=$sm timeout 300 python scripts/aas_bench.py 4194304 2>&1 | grep "TraceClientBBox x"; done
