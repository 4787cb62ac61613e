#!/bin/bash
# last check of the final build: smoke, full GPU suite, C2 (all legs, kernel table) and C4b / C1 / C3 bench lines, ncu launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/kernels_b32.json > gpurun_out/bench_C2_n1.json 2> gpurun_out/bench_C2_n1.err
for cfg in C1 C3 C4 C4b; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_${cfg}_n1_final.json 2> gpurun_out/bench_${cfg}_n1.err
done
timeout 300 python bench.py --quick --steps 2 --warmup 3 > gpurun_out/quick.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_quick.log 2>&1
for f in gpurun_out/bench_C2_n1.json gpurun_out/bench_*_n1_final.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "clocks")}, "e2e", d.get("e2e", {}).get("value"), "roof", (d.get("roofline") or {}).get("frac"),
          "lib", (d.get("gpu_library_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline_t8") or {}).get("value"))
except Exception as e:
    print("FAILED", e)
PY
done
