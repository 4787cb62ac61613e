#!/bin/bash
# ncu --set full of every non-conv kernel of one C2 step (memory-bound blocks, attention cores, NMS, stem, fused Swin), plus the conv kernel's
# first launches; summaries are extracted locally with tools/ncu_summary.py
mkdir -p gpurun_out
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"gap_partial|cbam_|ca_|scale_channels|sppf|nms_|stem_mma|se_gate|dwconv|layernorm|window_|mha_|bilinear|adaptive|scale_boxes|scale_weights|swin64" \
  -s 70 -c 40 -o gpurun_out/prof_blocks python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_blocks.log 2>&1
tail -n 3 gpurun_out/ncu_blocks.log; ls -la gpurun_out/prof_blocks.ncu-rep
