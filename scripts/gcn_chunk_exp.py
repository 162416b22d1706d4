"""GPU experiment: ST-GCN GEMM chain time and error against the fp32 FFMA chain for different TMEM drain intervals."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
h = P.make_handle(with_imu=False)
g = P.golden("gcn2.npz")
B = 2048
torch.manual_seed(0)
x = (torch.randn(B, 3, 20, 15, 1, device="cuda") * 0.5).contiguous()
h.set_option("gcn_gemm", 0)
ref = h.gcn_extract_feature(x).double()
h.set_option("gcn_gemm", 1)
for ch in (4, 8, 0, 2):
    h.set_option("gcn_kb_chunk", ch)
    for _ in range(2):
        out = h.gcn_extract_feature(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = h.gcn_extract_feature(x)
    e1.record()
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"gcn_kb_chunk {ch}: {e0.elapsed_time(e1) / 5:.3f} ms per 2048 snippets, rel max err vs FFMA {err:.2e}")
