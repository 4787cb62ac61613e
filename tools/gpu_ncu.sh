#!/bin/bash
# ncu --set full captures of the tcgen05 conv kernel on representative layer shapes (args of tools/prof_conv.py)
mkdir -p gpurun_out
i=0
for shape in ${SHAPES:-"64 160 160 64 64 1 1 5 1792" "64 160 160 64 64 1 1 5 0"}; do
  i=$((i+1))
  timeout 120 python tools/prof_conv.py $shape > gpurun_out/prof_plain_$i.log 2>&1 && \
  timeout 600 ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv_$i python tools/prof_conv.py $shape > gpurun_out/ncu_conv_$i.log 2>&1
  cat gpurun_out/prof_plain_$i.log; tail -n 2 gpurun_out/ncu_conv_$i.log
done
