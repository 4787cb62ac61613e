#!/bin/bash
# GPU tests + bench with per-op event table (no ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== tests"; tail -n 12 gpurun_out/t_gpu.log
echo "== bench"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
