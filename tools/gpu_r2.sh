#!/bin/bash
# one gpurun call (round 2): GPU tests, C2 bench (+ per-kernel event table), ncu launch list of the quick bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x --durations=15 2>&1 | tail -n 45 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" != "nocu" ]; then
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
fi
echo "== tests"; tail -n 30 gpurun_out/t_gpu.log
echo "== bench"; cut -c1-3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "== ncu"; tail -n 3 gpurun_out/ncu_quick.log; wc -l gpurun_out/launches.csv
