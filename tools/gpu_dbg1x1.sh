#!/bin/bash
# what bounds the deep 1x1 convs: normal / epilogue drains without work (1) / producer skips TMA (2) / issuer skips MMA (4) / no TMA store (8)
for shape in "32 40 40 256 256 1 1" "32 40 40 512 256 1 1" "32 20 20 512 512 1 1" "32 40 40 256 768 1 1" "32 80 80 256 128 1 1"; do
  for dbg in 0 1 2 4 8; do PROF_WARM=1 timeout 60 python tools/prof_conv.py $shape 9 $((dbg*256)) 2>&1 | tail -n 1 | cut -c1-200 | sed "s/^/dbg$dbg /"; done
done
