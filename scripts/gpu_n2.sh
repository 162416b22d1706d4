#!/bin/bash
# quick 2-GPU check of the bench (weak + strong) after kernel changes
mkdir -p gpurun_out
for MODE in weak strong; do
  OUT=gpurun_out/n2_${MODE}.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 6 --warmup 3 --scaling $MODE --no-cpu-baseline --no-config1 --no-half > $OUT 2> ${OUT%.json}.err
  echo "N=2 $MODE rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT"))
print("$MODE", round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "coll", d.get("collective_ms_per_step"), "launches", d["gpu_launches"])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -c 400
