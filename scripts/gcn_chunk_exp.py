"""GPU experiment: ST-GCN GEMM chain time and error against the fp32 FFMA chain (test-only variant library) for different
TMEM drain intervals (gcn_kb_chunk; 0 = accumulate a whole tile in TMEM)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
from mmego_b200 import _capi, build as B_
h = P.make_handle(lib=_capi.Lib(B_.VARIANTS["ffma"]["lib"]), with_imu=False)
B = 2048
torch.manual_seed(0)
x = (torch.randn(B, 3, 20, 15, 1, device="cuda") * 0.5).contiguous()
h.set_option("gcn_gemm", 0)
ref = h.gcn_extract_feature(x).double()
h.set_option("gcn_gemm", 1)
up_sd, lo_sd = P.checkpoints()
for ch in (4, 8, 0):
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
    errs = []
    for name in P.SWEEP_GOLDENS[:1] + ("synth3.npz",):
        g = P.golden(name)
        Bq, L = g["data"].shape[:2]
        xq = g["data"].cuda().clone()
        h0 = torch.zeros(6, Bq, 64, device="cuda")
        l = h.upper_forward(xq, h0, h0.clone(), g["skl"].cuda(), g["R"].cuda(), g["t"].cuda())[0]
        ll, ql = h.lower_forward(g["upper_l"].cuda(), xq, g["skl"].cuda(), g["R"].cuda(), g["t"].cuda())
        errs.append((P.maxerr(ll, g["lower_l"]), P.rot_angle_deg(ql, g["q_lower"])))
    print(f"gcn_kb_chunk {ch}: {e0.elapsed_time(e1) / 5:.3f} ms per 2048 snippets, rel max err vs FFMA {err:.2e}; lower joints / q_lower vs reference goldens {errs}")
