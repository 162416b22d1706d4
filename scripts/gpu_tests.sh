#!/bin/bash
# parity suite only.   usage: gpu_tests.sh <tag> [pytest -k expression]
mkdir -p gpurun_out
TAG=${1:-t}; shift
rm -f gpurun_out/parity_ragged_fp64.jsonl
if [ -n "$1" ]; then K=(-k "$1"); else K=(); fi
timeout 1500 python -m pytest tests -m gpu -q -s "${K[@]}" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_${TAG}.log
