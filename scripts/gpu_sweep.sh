#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "infer_host or pin" 2>&1 | tail -5 || { echo "TEST FAILED"; exit 1; }
timeout 300 python bench.py --steps 5 --warmup 3 --no-half > gpurun_out/bench.json 2> gpurun_out/bench.err || { echo BENCH FAILED; tail gpurun_out/bench.err; exit 1; }
python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
P
timeout 600 python scripts/sweep.py 2>&1 | tail -14
