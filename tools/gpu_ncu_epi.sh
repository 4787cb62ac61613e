#!/bin/bash
# where do the conv kernel's warps stall on epilogue-dominated shapes? ncu --set full with source, top SASS lines by stall samples
mkdir -p gpurun_out
i=0
for sh in "12 160 160 64 64 1 1" "32 160 160 32 32 3 1" "32 160 160 64 128 3 1"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_epi_$i python tools/prof_conv.py $sh 5 > gpurun_out/ncu_epi_$i.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_epi_$i.ncu-rep --stalls 40 > gpurun_out/prof_epi_$i.txt 2>&1; rm -f gpurun_out/prof_epi_$i.ncu-rep
  tail -n 1 gpurun_out/ncu_epi_$i.log
done
