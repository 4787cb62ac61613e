#!/bin/bash
# fused P2 SwinBlock on tcgen05: unit test against the mma.sync kernel, event timings of both, model parity of the layer
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_swin.py -q -m gpu -x 2>&1 | tail -n 6
timeout 120 python tools/prof_swin.py 32 160 160 9 2>&1 | tail -n 2
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "swin or sod_bf16_640 or fp16" 2>&1 | tail -n 4
