#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_conv.py 32 160 160 64 64 1 1 5 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv1x1 python tools/prof_conv.py 32 160 160 64 64 1 1 5 > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv3x3 python tools/prof_conv.py 32 160 160 64 64 3 1 5 >> gpurun_out/ncu_conv.log 2>&1
tail -n 4 gpurun_out/ncu_conv.log
