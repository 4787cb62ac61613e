#!/bin/bash
# role isolation of the tcgen05 conv kernel: debug bits (mode>>8): 1 no epilogue work, 2 no TMA loads, 4 no MMA
mkdir -p gpurun_out; rm -f gpurun_out/roles.txt
for shape in "32 160 160 64 64 3 1 7" "32 160 160 64 128 3 2 7" "32 160 160 64 64 1 1 7" "32 160 160 32 32 3 1 7" "32 40 40 128 128 3 1 7" "32 20 20 256 256 3 1 7" "32 80 80 256 128 1 1 7"; do
  for dbg in 0 1 2 4 3 6 5 7; do
    echo -n "dbg=$dbg " >> gpurun_out/roles.txt
    timeout 120 python tools/prof_conv.py $shape $((dbg*256)) silu >> gpurun_out/roles.txt 2>&1
  done
done
echo -n "gelu dbg=0 " >> gpurun_out/roles.txt; timeout 120 python tools/prof_conv.py 1 1 829472 64 128 1 1 7 0 gelu >> gpurun_out/roles.txt 2>&1
echo -n "none dbg=0 " >> gpurun_out/roles.txt; timeout 120 python tools/prof_conv.py 1 1 829472 64 128 1 1 7 0 none >> gpurun_out/roles.txt 2>&1
echo -n "silu dbg=0 " >> gpurun_out/roles.txt; timeout 120 python tools/prof_conv.py 1 1 829472 64 128 1 1 7 0 silu >> gpurun_out/roles.txt 2>&1
cat gpurun_out/roles.txt
