#!/bin/bash
# usage: scripts/prof3.sh <tag> <python script + args> -- <kernel func substr> <source file>
set -e
tag=$1; shift
gpurun --timeout 1200 -- "python $1 $2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trace -s 1 -c 1 -f -o gpurun_out/prof_$tag python $1 $2 > gpurun_out/ncu.log 2>&1; tail -n 2 gpurun_out/ncu.log" 2>&1 | tail -4
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv 2>/dev/null > /tmp/raw_$tag.csv
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv 2>/dev/null > /tmp/src_$tag.csv
rm -rf /tmp/cub_$tag && mkdir /tmp/cub_$tag && (cd /tmp/cub_$tag && cuobjdump -xelf all /root/repo/threecore_b200/libcmgpu.so >/dev/null 2>&1 && nvdisasm -g -c cm_gpu.sm_100a.cubin > dis.txt 2>&1)
python scripts/ncu_summary.py /tmp/raw_$tag.csv | grep -v "stalled_\(barrier\|drain\|lg_thr\|membar\|misc\|sleeping\|tex_thr\|dispatch\|mio\)"
python scripts/ncu_blocks.py /tmp/src_$tag.csv /tmp/cub_$tag/dis.txt $3 $4
