#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/conv_shapes.txt
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -n 40 > gpurun_out/t_conv.log
for shape in "32 160 160 64 64 3 1 7 1" "32 160 160 64 64 3 1 7 2" "32 80 80 64 64 3 1 7 1" "32 80 80 64 64 3 1 7 2" "32 40 40 128 128 3 1 7 1" "32 40 40 128 128 3 1 7 2" "32 80 80 128 64 3 1 7 1" "32 80 80 128 64 3 1 7 2"; do
  timeout 120 python tools/prof_conv.py $shape >> gpurun_out/conv_shapes.txt 2>&1
done
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -n 60 > gpurun_out/t_model.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== conv"; tail -n 12 gpurun_out/t_conv.log
echo "== model"; tail -n 8 gpurun_out/t_model.log
echo "== bench"; cat gpurun_out/bench.json | cut -c1-300; tail -n 3 gpurun_out/bench.err
cat gpurun_out/conv_shapes.txt
