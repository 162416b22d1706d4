"""Per-stage device split of ONE one-snippet step (library profile spans; eager calls, so the spans include launch gaps)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmego_b200.Processor.Test.Demo_test import MMEgo
m = MMEgo(batch_size=1, imu_surrogate=False, quiet=True, use_graph=False)
for f in ("data", "target", "skl", "imu"):
    setattr(m, f, getattr(m, f)[:200])
m.eval_model()
h = m.pipe.handle
h.profile_begin()
m.eval_model()
prof = h.profile_read()
rec = {k: round(v["ms"] * 1e3 / v["spans"], 2) for k, v in prof.items() if not k.startswith("gcn.")}
print("eager, us per call:", rec, "sum", round(sum(rec.values()), 1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "lat_breakdown_final.json"), "w"), indent=1)
