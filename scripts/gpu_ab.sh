#!/bin/bash
# A/B of library options on the bench: scripts/gpu_ab.sh "<opts A>" "<opts B>" ...
mkdir -p gpurun_out
i=0
for o in "$@"; do
  i=$((i+1))
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-half $o > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab$i.json"))
print("[$o]", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]) if "e2e" in d else None, "clk", d["clocks"]["sm_mhz"], {k:v for k,v in d["stage_ms_per_step"].items() if k.startswith("imu.lstm")})
PY
done
