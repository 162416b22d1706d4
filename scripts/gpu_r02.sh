#!/bin/bash
# One GPU-box visit (round 2): parity suite, bench (+ reference arm), logs under gpurun_out/.   usage: gpu_r02.sh <tag> [bench args]
mkdir -p gpurun_out
TAG=${1:-r02}; shift
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu_${TAG}.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, "clk", d["clocks"]["sm_mhz"])
    print(d["stage_ms_per_step"])
    if d.get("half_mode"): print("half", round(d["half_mode"]["value"]), round(d["half_mode"]["ms_per_step"],2))
    print("config1", d.get("config1"))
    print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("kind"))
    r=json.load(open("gpurun_out/bench_ref_${TAG}.json")); print("ref", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
except Exception as e:
    print("summary failed:", e)
PY
