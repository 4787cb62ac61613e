#!/bin/bash
# GPU tests + bench + ncu --set full of the memory-bound kernels of one step (stem, SE, CBAM, CA, SPPF, upsample, decode, dw)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_mma|dfl_decode|gap_partial|scale_channels|cbam_apply|cbam_stats|upsample|sppf|dwconv|ca_pool|ca_apply|layernorm_stream|mha_" -s 66 -c 33 -o gpurun_out/prof_mem python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_mem.log 2>&1
echo "== tests"; tail -n 8 gpurun_out/t_gpu.log
echo "== bench"; cut -c1-300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
tail -n 2 gpurun_out/ncu_mem.log; ls -la gpurun_out/prof_mem.ncu-rep
