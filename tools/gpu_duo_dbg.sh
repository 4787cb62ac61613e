#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -n 5
python tools/trace_conv.py 32 160 160 32 32 3 1 0 > gpurun_out/trace_duo.txt 2>&1
python tools/trace_conv.py 32 160 160 32 32 3 1 256 > gpurun_out/trace_duo_noepi.txt 2>&1
tail -n 14 gpurun_out/trace_duo.txt | cut -c1-330
for dbg in 0 1 8 2 4; do timeout 60 python tools/prof_conv.py 32 160 160 32 32 3 1 7 $((dbg*256)) 2>&1 | tail -n 1 | cut -c1-220; done
