#!/bin/bash
# quick fail-fast check of selected tests, then the full round
set -o pipefail
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "test_gcn" 2>&1 | tail -15 || { echo "GCN TEST FAILED"; exit 1; }
bash scripts/gpu_round.sh
