#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -n 60 > gpurun_out/t_model.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
for shape in "32 160 160 64 64 3 1" "32 80 80 64 64 3 1" "32 40 40 128 128 3 1" "32 160 160 32 32 3 1" "32 20 20 256 256 3 1" "32 80 80 256 128 1 1" "32 160 160 64 64 1 1" "32 160 160 64 128 3 2" "32 20 20 512 512 1 1"; do
  timeout 120 python tools/prof_conv.py $shape 7 >> gpurun_out/conv_shapes.txt 2>&1
done
timeout 300 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/quick.json 2>gpurun_out/quick.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 272 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_quick.log 2>&1
timeout 120 python tools/prof_conv.py 32 160 160 64 64 3 1 5 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv64 python tools/prof_conv.py 32 160 160 64 64 3 1 5 > gpurun_out/ncu_conv.log 2>&1
echo "== model"; tail -n 4 gpurun_out/t_model.log
echo "== bench"; cat gpurun_out/bench.json | cut -c1-600; tail -n 3 gpurun_out/bench.err
cat gpurun_out/conv_shapes.txt
tail -n 3 gpurun_out/ncu_quick.log gpurun_out/ncu_conv.log
