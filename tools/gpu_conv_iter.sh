#!/bin/bash
# conv kernel iteration: correctness (conv + model tests), role isolation on key shapes, traces, bench
mkdir -p gpurun_out; rm -f gpurun_out/roles.txt
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 15 > gpurun_out/t_conv.log
cat gpurun_out/t_conv.log
for shape in "32 160 160 64 64 3 1 7" "32 160 160 64 128 3 2 7" "32 160 160 64 64 1 1 7" "32 160 160 32 32 3 1 7" "32 40 40 128 128 3 1 7" "32 80 80 256 128 1 1 7"; do
  for dbg in ${DBGS:-0 3 6 7}; do
    echo -n "dbg=$dbg " >> gpurun_out/roles.txt
    timeout 120 python tools/prof_conv.py $shape $((dbg*256)) silu >> gpurun_out/roles.txt 2>&1
  done
done
echo -n "gelu dbg=0 " >> gpurun_out/roles.txt; timeout 120 python tools/prof_conv.py 1 1 829472 64 128 1 1 7 0 gelu >> gpurun_out/roles.txt 2>&1
cat gpurun_out/roles.txt | cut -c1-190
python tools/trace_conv.py 32 160 160 64 64 3 1 0 > gpurun_out/trace_halo.txt 2>&1
python tools/trace_conv.py 32 160 160 64 64 1 1 0 > gpurun_out/trace_1x1.txt 2>&1
python tools/trace_conv.py 32 160 160 64 128 3 2 0 > gpurun_out/trace_s2.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
