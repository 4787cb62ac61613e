#!/bin/bash
# Round evidence in one gpurun call: GPU tests, full bench line (+ per-op event table), reference arm, config table (C1-C5),
# ncu launch list (time + DRAM bytes) of the quick bench, ncu --set full of the fused Swin kernel, the stem and four conv launches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 15 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"swin64_fused|stem_mma" -s 6 -c 2 -o gpurun_out/prof_swin_stem python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 245 -c 4 -o gpurun_out/prof_conv_step python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
echo "== tests"; tail -n 6 gpurun_out/t_gpu.log
echo "== bench"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "== ref"; cut -c1-300 gpurun_out/bench_ref.json
echo "== configs"; cut -c1-600 gpurun_out/configs.json; tail -n 2 gpurun_out/configs.err
tail -n 2 gpurun_out/ncu_quick.log gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log; ls -la gpurun_out/*.ncu-rep
