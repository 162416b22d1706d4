#!/bin/bash
mkdir -p gpurun_out
for hc in 512 1024 2048; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-half --opt host_chunk=$hc > gpurun_out/bench_hc$hc.json 2> gpurun_out/bench_hc$hc.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_hc$hc.json"))
print("host_chunk", $hc, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"])
PY
done
