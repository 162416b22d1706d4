#!/bin/bash
# A/B of two prebuilt libraries inside one box visit: LIB_A / LIB_B are .so paths copied over the product library
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "imu" 2>&1 | tail -3
for i in 1 2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config1 --no-half > gpurun_out/bench_pool_$i.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pool_$i.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k:v for k,v in d["stage_ms_per_step"].items() if not k.startswith("gcn.")}, d["stages"]["imu.pool"])
PY
done
