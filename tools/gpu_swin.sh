#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_swin.py -q -m gpu -x 2>&1 | tail -n 25
