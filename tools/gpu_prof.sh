#!/bin/bash
# Evidence for profiles/: bench line, per-kernel event table, ncu launch list (time + DRAM bytes per launch) of one quick step,
# ncu --set full of the fused Swin kernel and of four tcgen05 conv launches. Keep reports small: gpurun copies back <= 64 MiB.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin64_fused -s 3 -c 1 -o gpurun_out/prof_swin64 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 252 -c 4 -o gpurun_out/prof_conv_step python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | head -30
cut -c1-300 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err; tail -n 2 gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log
