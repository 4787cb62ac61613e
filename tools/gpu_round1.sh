#!/bin/bash
# first GPU trip: each suite in its own process so one faulting kernel cannot take the others down
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_nms.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_nms.log
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -80 > gpurun_out/t_conv.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -120 > gpurun_out/t_model.log
tail -5 gpurun_out/t_nms.log gpurun_out/t_conv.log gpurun_out/t_model.log
