"""Error of the point-encoder outputs (upper.g, gw, lower.ak) of point_gemm=0 (FFMA) and 1 (mma.sync fp16x3) against
an fp64 evaluation of the oracle.  Runs on the emulated library (CPU) or, with --gpu, on the product library."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmego_b200 import _capi
from tests import _parity as P
from oracle import mmego_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--gpu", action="store_true")
ap.add_argument("--n", type=int, default=2)
a = ap.parse_args()
if a.gpu:
    h = P.make_handle(with_imu=False)
else:
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul
    h = P.make_handle(lib=_capi.Lib(build_emul.build()), require_cuda=False, with_imu=False)
g = P.golden("sample16.npz")
up_sd, lo_sd = P.checkpoints()
s = slice(0, a.n)
b = a.n
ref = {}
for mode in (0, 1):
    h.set_option("point_gemm", mode)
    x = P.dev(h, g["data"][s].clone())
    h0 = torch.zeros(6, b, 64, device=h.device)
    skl, R, t = P.dev(h, g["skl"][s]), P.dev(h, g["R"][s]), P.dev(h, g["t"][s])
    gt = torch.zeros(b * 20, 64, device=h.device)
    h.tap("upper.g", gt)
    l, q, gw, hn, cn = h.upper_forward(x, h0, h0.clone(), skl, R, t)
    x = P.dev(h, g["x1"][s].clone())
    ak = torch.zeros(b * 20, 192, device=h.device)
    h.tap("lower.ak", ak)
    ll, ql = h.lower_forward(P.dev(h, g["upper_l"][s]), x, skl, R, t)
    out = dict(g=gt.cpu(), gw=gw.cpu(), upper_l=l.cpu(), ak=ak.cpu(), lower_l=ll.cpu())
    if mode == 0:
        ref = out
    print("mode", mode, "upper_l vs golden %.3e" % P.maxerr(l, g["upper_l"][s]), "gw vs golden %.3e" % P.maxerr(gw.reshape(b, 20, -1), g["gw"][s]),
          "lower_l vs golden %.3e" % P.maxerr(ll, g["lower_l"][s]))
    if mode == 1:
        for k in out:
            print("  mma vs ffma  %-8s max|d| %.3e   (max|ref| %.3e)" % (k, P.maxerr(out[k], ref[k]), float(ref[k].abs().max())))
