#!/bin/bash
# Round-2 evidence in one call: full GPU test suite, C2 bench (all baselines, per-kernel table), the other BASELINE configs at N=1 with
# the clock sampler, ncu launch list (+ DRAM bytes) of a quick C2 step.
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -n 20 > gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench_C2_n1.json 2> gpurun_out/bench_C2_n1.err
for cfg in C1 C3 C4 C4b C5u C5c; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/bench_${cfg}_n1.json 2> gpurun_out/bench_${cfg}_n1.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
echo "== tests"; tail -n 12 gpurun_out/t_gpu.log
for f in gpurun_out/bench_*_n1.json gpurun_out/bench_reference_arm.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "clocks")}, "e2e", d.get("e2e", {}).get("value"), "roof", (d.get("roofline") or {}).get("frac"),
          "lib", (d.get("gpu_library_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline_t8") or {}).get("value"))
except Exception as e:
    print("FAILED", e)
PY
done
wc -l gpurun_out/launches.csv
