#!/bin/bash
# GPU round: parity tests, smoke, bench (fail-fast: a hung kernel must not burn the budget)
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log || { echo "PYTEST FAILED"; exit 1; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log || { echo "SMOKE FAILED"; exit 1; }
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err || { echo "BENCH FAILED"; tail -20 gpurun_out/bench.err; exit 1; }
cat gpurun_out/bench.json
