#!/bin/bash
# A/B/... of library option SETS inside ONE box visit.  usage: gpu_abn.sh "<k=v,k=v>" "<k=v>" ...   ("-" = defaults)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "imu" 2>&1 | tail -3
i=0
for round in 1 2; do
for SET in "$@"; do
  i=$((i+1))
  OPTS=""
  if [ "$SET" != "-" ]; then for kv in ${SET//,/ }; do OPTS="$OPTS --opt $kv"; done; fi
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-config1 --no-half $OPTS > gpurun_out/abn_$i.json 2> gpurun_out/abn.err || tail -3 gpurun_out/abn.err
  python - <<PY
import json
d=json.load(open("gpurun_out/abn_$i.json"))
print("$SET", "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k:v for k,v in d["stage_ms_per_step"].items() if not k.startswith("gcn.")})
PY
done
done 2>&1 | tee gpurun_out/abn.log
