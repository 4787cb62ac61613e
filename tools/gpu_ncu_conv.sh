#!/bin/bash
# ncu --set full of conv_tc_kernel: (a) five representative layer shapes in isolation (tools/prof_conv.py), (b) the conv launches of one C2 step
# that carry the round-2 plans (back-to-back GEMM heads / C2f, tile pairs, flat 1x1)
mkdir -p gpurun_out
i=0
for sh in "32 160 160 64 64 3 1" "32 160 160 64 128 3 1" "32 40 40 128 128 3 1" "32 20 20 256 256 3 1" "32 160 160 128 64 1 1" "32 160 160 32 32 3 1" "32 40 40 512 256 1 1"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_conv_$i python tools/prof_conv.py $sh 5 > gpurun_out/ncu_conv_$i.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_conv_$i.ncu-rep > gpurun_out/prof_conv_$i.txt 2>&1; rm -f gpurun_out/prof_conv_$i.ncu-rep
  tail -n 1 gpurun_out/ncu_conv_$i.log
done
timeout 300 python bench.py --quick --steps 2 --warmup 3 --no-overlap > gpurun_out/quick.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 150 -c 75 --import-source off -o gpurun_out/prof_conv_step python bench.py --quick --steps 2 --warmup 3 --no-overlap > gpurun_out/ncu_conv_step.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_conv_step.ncu-rep > gpurun_out/prof_conv_step.txt 2>&1; rm -f gpurun_out/prof_conv_step.ncu-rep
tail -n 2 gpurun_out/ncu_conv_step.log; du -sh gpurun_out
