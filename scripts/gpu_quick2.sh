#!/bin/bash
# quick iteration: a few parity tests + a short device-only bench.   usage: gpu_quick2.sh <tag> [pytest -k expr] [bench args...]
mkdir -p gpurun_out
TAG=${1:-q}; KEXPR=${2:-"synth3 or sample16 or sweep or config_shape"}; shift; shift
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config1 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"])
    print(d["stage_ms_per_step"])
    if d.get("half_mode"): print("half", round(d["half_mode"]["value"]), round(d["half_mode"]["ms_per_step"],2), d["half_mode"]["stage_ms_per_step"])
except Exception as e:
    print("summary failed:", e); print(open("gpurun_out/bench_${TAG}.err").read()[-2000:])
PY
