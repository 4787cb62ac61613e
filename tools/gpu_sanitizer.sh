#!/bin/bash
# compute-sanitizer passes over smoke() and the tcgen05 conv tests (hand-rolled mbarrier / TMEM pipeline): memcheck, then racecheck
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" | tee gpurun_out/san_rc.txt
timeout 1200 $S --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "conv_tc and not fused_upsample" > gpurun_out/san_memcheck_conv.log 2>&1; echo "memcheck conv rc=$?" | tee -a gpurun_out/san_rc.txt
timeout 1200 $S --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" | tee -a gpurun_out/san_rc.txt
timeout 600 $S --tool synccheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?" | tee -a gpurun_out/san_rc.txt
for f in gpurun_out/san_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|passed|failed|Error|hazard" $f | sort | uniq -c | head -12; done
