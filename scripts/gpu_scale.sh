#!/bin/bash
# 1 -> 8 GPU scaling curves on ONE box: weak (B=4096 per GPU) and strong (B=4096 in total) at N = 1, 2, 4, 8.
mkdir -p gpurun_out
TAG=${1:-r02}
for N in 1 2 4 8; do
  for MODE in weak strong; do
    if [ $N -eq 1 ] && [ $MODE = strong ]; then continue; fi
    OUT=gpurun_out/scale_${TAG}_n${N}_${MODE}.json
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-config1 --no-half > $OUT 2> ${OUT%.json}.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --steps 10 --warmup 3 --scaling $MODE --no-cpu-baseline --no-config1 --no-half > $OUT 2> ${OUT%.json}.err
    fi
    echo "N=$N $MODE rc=$?"
  done
done
python - <<PY
import json, glob
base=None
for N in (1,2,4,8):
    for mode in ("weak","strong"):
        try:
            d=json.load(open(f"gpurun_out/scale_${TAG}_n{N}_{mode}.json"))
        except Exception as e:
            continue
        if N==1: base=d["value"]
        print(f"N={N} {mode:6s} value {d['value']:.0f} ms/step {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.0f} eff {d['value']/(base*N) if base else 0:.3f} coll_ms {d.get('collective_ms_per_step')} lstm_fast {d['stage_ms_per_step'].get('imu.lstm_fast')} lstm_slow {d['stage_ms_per_step'].get('imu.lstm_slow')}")
PY
