#!/bin/bash
# epilogue cost components on an epilogue-dominated, L2-resident shape: 64 -> 64 1x1 @160^2 x 12 images (4 MMAs per 128 x 64 tile)
for act in silu none gelu relu; do
  for dbg in 0 16 8 24 1; do
    echo -n "act=$act dbg=$dbg: "; PROF_WARM=1 timeout 60 python tools/prof_conv.py 12 160 160 64 64 1 1 15 $((dbg*256)) $act 2>&1 | tail -n 1 | cut -c60-200
  done
done
echo "== fp32 out"; 
