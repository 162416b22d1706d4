"""GPU experiment: ST-GCN chain time and error against the fp32 FFMA chain for gcn_gemm = 1 (15-row frames, nine shifted
tile loads per temporal conv) and 2 (16-row frames, one resident window per channel block)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
h = P.make_handle(with_imu=False)
for B, T in ((2, 20), (3, 7), (2048, 20)):
    torch.manual_seed(0)
    x = (torch.randn(B, 3, T, 15, 1, device="cuda") * 0.5).contiguous()
    h.set_option("gcn_gemm", 0)
    ref = h.gcn_extract_feature(x).double()
    for mode in (1, 2):
        h.set_option("gcn_gemm", mode)
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
        print(f"B={B} T={T} gcn_gemm {mode}: {e0.elapsed_time(e1) / 5:.3f} ms, rel max err vs FFMA {err:.2e}", flush=True)
