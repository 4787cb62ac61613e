#!/bin/bash
# isolated timing (L2 flushed) + CTA-0 pipeline traces of the deep / wide conv shapes of SOD-simple at batch 32
mkdir -p gpurun_out
{
for sh in "32 40 40 128 128 3 1" "32 20 20 256 256 3 1" "32 40 40 512 256 1 1" "32 80 80 64 128 3 2" "32 40 40 128 256 3 2" "32 20 20 1024 512 1 1" \
          "32 80 80 256 128 1 1" "32 40 40 256 128 3 1" "32 80 80 128 128 3 1" "32 20 20 512 128 3 1" "32 80 80 64 64 3 1" "32 160 160 32 32 3 1" \
          "32 160 160 64 64 1 1" "32 160 160 96 64 1 1" "32 40 40 384 256 1 1" "32 20 20 768 512 1 1" "32 40 40 256 256 1 1"; do
  timeout 120 python tools/prof_conv.py $sh 7 $1
done
} > gpurun_out/conv_shapes.txt 2>&1
timeout 120 python tools/trace_conv.py 32 40 40 128 128 3 1 > gpurun_out/trace_128_3x3_40.txt 2>&1
timeout 120 python tools/trace_conv.py 32 20 20 256 256 3 1 > gpurun_out/trace_256_3x3_20.txt 2>&1
timeout 120 python tools/trace_conv.py 32 40 40 512 256 1 1 > gpurun_out/trace_512_256_1x1_40.txt 2>&1
timeout 120 python tools/trace_conv.py 32 80 80 64 128 3 2 > gpurun_out/trace_64_128_s2_80.txt 2>&1
cat gpurun_out/conv_shapes.txt
