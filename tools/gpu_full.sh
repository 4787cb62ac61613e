#!/bin/bash
# one gpurun call: GPU tests, bench (+ per-kernel event table), ncu launch list of the quick bench, ncu --set full of the top kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 30 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|swin64_fused|stem_mma" -s 267 -c 89 -o gpurun_out/prof_step_conv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "== tests"; tail -n 8 gpurun_out/t_gpu.log
echo "== bench"; cut -c1-1800 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "== ncu"; tail -n 3 gpurun_out/ncu_quick.log; wc -l gpurun_out/launches.csv; tail -n 2 gpurun_out/ncu_full.log
