#!/bin/bash
# End-of-session evidence: GPU tests, smoke(), full bench line (+ per-op table), reference arm, config table, ncu launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 15 > gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
echo "== tests"; tail -n 4 gpurun_out/t_gpu.log; cat gpurun_out/smoke.log | tail -2
echo "== bench"; cut -c1-300 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
echo "== ref"; cut -c1-200 gpurun_out/bench_ref.json
