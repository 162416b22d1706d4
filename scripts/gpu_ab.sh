#!/bin/bash
# A/B of a library option inside ONE box visit.   usage: gpu_ab.sh <opt> <valA> <valB> [bench args]
OPT=$1; A=$2; B=$3; shift; shift; shift
mkdir -p gpurun_out
for V in $A $B $A $B; do
  python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config1 --no-half --opt $OPT=$V "$@" > gpurun_out/ab_${OPT}_$V.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${OPT}_$V.json"))
print("$OPT=$V", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), {k:v for k,v in d["stage_ms_per_step"].items() if not k.startswith("gcn.")})
PY
done
