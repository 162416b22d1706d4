#!/bin/bash
# N-GPU runs (weak and strong scaling) through torchrun.   usage: gpu_multi2.sh <N> <tag>
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-m}
for MODE in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --scaling $MODE --no-cpu-baseline --no-config1 \
    > gpurun_out/bench_n${N}_${MODE}_${TAG}.json 2> gpurun_out/bench_n${N}_${MODE}_${TAG}.err; echo "$MODE rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_${MODE}_${TAG}.json"))
    print("$MODE N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, "coll_ms", d.get("collective_ms_per_step"), d["config"]["workload"][:60])
    print(d["stage_ms_per_step"])
except Exception as e:
    print("failed:", e); print(open("gpurun_out/bench_n${N}_${MODE}_${TAG}.err").read()[-1500:])
PY
done
