#!/bin/bash
# A/B evidence for the fused P2 SwinBlock: mma.sync kernel vs tcgen05 / TMEM kernel -- event timings, then ncu --set full with source
mkdir -p gpurun_out
timeout 120 python tools/prof_swin.py 32 160 160 7 2>&1 | tail -n 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin64 -s 2 -c 2 -o gpurun_out/prof_swin python tools/prof_swin.py 32 160 160 2 > gpurun_out/ncu_swin.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_swin.ncu-rep --stalls 45 > gpurun_out/prof_swin.txt 2>&1; rm -f gpurun_out/prof_swin.ncu-rep
tail -n 2 gpurun_out/ncu_swin.log
