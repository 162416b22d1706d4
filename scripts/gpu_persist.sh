#!/bin/bash
# persistent-timestep LSTM launches: bit-identity tests (bounded by timeout), then an A/B inside one box visit
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "persistent_timesteps" 2>&1 | tail -5
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
for V in ${@:-0 1 3 0 1 3}; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config1 --no-half --opt tc_persist=$V > gpurun_out/ab_tc_persist_${V}.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_tc_persist_${V}.json"))
print("tc_persist=$V", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k:v for k,v in d["stage_ms_per_step"].items() if not k.startswith("gcn.")}, "launches", d.get("gpu_launches"))
PY
done
