#!/bin/bash
# ncu --set full of the final fused SwinBlock kernels (mma.sync baseline + tcgen05) and of the stem (fp32 / uint8 source, with pooling partials)
mkdir -p gpurun_out
timeout 120 python tools/prof_swin.py 32 160 160 5 2>&1 | tail -n 12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin64 -s 2 -c 2 -o gpurun_out/prof_swin python tools/prof_swin.py 32 160 160 2 > gpurun_out/ncu_swin.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_swin.ncu-rep --stalls 30 > gpurun_out/prof_swin.txt 2>&1; rm -f gpurun_out/prof_swin.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 4 -c 4 -o gpurun_out/prof_stem python tools/prof_stem.py 32 640 2 > gpurun_out/ncu_stem.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_stem.ncu-rep --stalls 20 > gpurun_out/prof_stem.txt 2>&1; rm -f gpurun_out/prof_stem.ncu-rep
tail -n 2 gpurun_out/ncu_swin.log gpurun_out/ncu_stem.log
