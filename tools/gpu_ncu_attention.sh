#!/bin/bash
# ncu --set full of the two attention implementations on the model shapes (tools/ab_attention.py launches impl 1 then impl 2 per shape)
mkdir -p gpurun_out
timeout 300 python tools/ab_attention.py > gpurun_out/ab_attention.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mha_tc_kernel|mha_win_kernel|mha_flash_kernel" -s 20 -c 6 \
   -o gpurun_out/prof_attention python tools/ab_attention.py > gpurun_out/ncu_attention.log 2>&1
tail -n 2 gpurun_out/ncu_attention.log
