#!/bin/bash
# Round-end verification on one B200: smoke(), full GPU suite, default bench run, reference arm.
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash scripts/gpu_r02.sh $TAG
