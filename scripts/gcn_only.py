"""Runs GCN.Model.extract_feature (mmego_gcn_extract_feature) alone on B snippets: ncu target for the ST-GCN kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h = P.make_handle(with_imu=False)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    h.set_option(k, int(v))
torch.manual_seed(0)
x = (torch.randn(B, 3, 20, 15, 1, device="cuda") * 0.5).contiguous()
for _ in range(3):
    out = h.gcn_extract_feature(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = h.gcn_extract_feature(x)
e1.record()
torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per call; error flag {h.debug_stats(reset=False)[7]}")
