#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "conv_tc" 2>&1 | tail -n 12 > gpurun_out/t_conv.log
cat gpurun_out/t_conv.log
for dbg in 0 3 6 7; do python tools/prof_conv.py 32 160 160 64 64 3 1 7 $((dbg*256)) silu; done
python tools/prof_conv.py 32 40 40 128 128 3 1 7 0 silu
python tools/prof_conv.py 32 80 80 64 64 3 1 7 0 silu
python tools/prof_conv.py 32 160 160 64 128 3 2 7 0 silu
python tools/prof_conv.py 32 160 160 64 64 1 1 7 0 silu
python tools/prof_conv.py 32 160 160 32 32 3 1 7 0 silu
python tools/prof_conv.py 32 80 80 256 128 1 1 7 0 silu
python tools/prof_conv.py 1 1 829472 64 128 1 1 7 0 gelu
python tools/trace_conv.py 32 160 160 64 64 3 1 0 > gpurun_out/trace_halo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 5
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
