"""Batch-1 evaluation (the reference's own setting, Demo_test.py:61): ms per snippet through MMEgo().eval_model(), the
device-side split per stage (library profile spans), and the host time per snippet when the GPU is not waited for."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmego_b200.Processor.Test.Demo_test import MMEgo

out = {}
for bs in (1, 4):
    m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True)
    m.eval_model()
    best = 1e9
    for _ in range(3):
        m.eval_model()
        best = min(best, m.seconds)
    n = m.data.shape[0]
    rec = dict(ms_per_snippet=best / n * 1e3, it_per_s=n / best)
    h = m.pipe.handle
    n0 = h.launch_count()
    h.profile_begin()
    m.eval_model()
    prof = h.profile_read()
    rec["launches_per_call"] = (h.launch_count() - n0) / ((n + bs - 1) // bs)
    rec["stage_us_per_call"] = {k: round(v["ms"] * 1e3 / v["spans"], 2) for k, v in prof.items() if not k.startswith("gcn.")}
    rec["stage_sum_us_per_call"] = round(sum(rec["stage_us_per_call"].values()), 1)
    out[f"batch{bs}"] = rec
    print(f"batch {bs}: {rec['ms_per_snippet']:.3f} ms per snippet ({rec['it_per_s']:.0f} it/s), "
          f"{rec['launches_per_call']:.0f} launches per call, stages sum {rec['stage_sum_us_per_call']} us per call")
    print("   ", rec["stage_us_per_call"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lat_breakdown.json"), "w"), indent=1)
