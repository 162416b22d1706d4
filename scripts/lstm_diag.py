"""GPU diagnostic: small-LSTM mma path vs FFMA path through the debug taps."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
h = P.make_handle(with_imu=False)
g = P.golden("synth3.npz")
b = 3
res = {}
for mode in (0, 1):
    h.set_option("small_lstm_gemm", mode)
    for null_state in (False, True):
        x = P.dev(h, g["data"].clone())
        h0 = None if null_state else torch.zeros(6, b, 64, device=h.device)
        c0 = None if null_state else torch.zeros(6, b, 64, device=h.device)
        skl, R, t = P.dev(h, g["skl"]), P.dev(h, g["R"]), P.dev(h, g["t"])
        tp = torch.zeros(b * 20, 128, device=h.device)
        h.tap("upper.lstm", tp)
        l = h.upper_forward(x, h0, c0, skl, R, t, want_state=not null_state)[0]
        torch.cuda.synchronize()
        res[("upper", mode, null_state)] = tp.cpu()
    x = P.dev(h, g["x1"].clone())
    tp = torch.zeros(b * 20, 128, device=h.device)
    ta = torch.zeros(b * 20, 192, device=h.device)
    h.tap("lower.lstm", tp)
    h.tap("lower.ak", ta)
    ll = h.lower_forward(P.dev(h, g["upper_l"]), x, skl, R, t)[0]
    torch.cuda.synchronize()
    res[("lower", mode)] = tp.cpu()
    res[("ak", mode)] = ta.cpu()
    print("mode", mode, "lower_l vs golden", P.maxerr(ll, g["lower_l"]))
print("upper state  : mma vs ffma", P.maxerr(res[("upper", 1, False)], res[("upper", 0, False)]))
print("upper null   : mma vs ffma", P.maxerr(res[("upper", 1, True)], res[("upper", 0, True)]))
print("lower        : mma vs ffma", P.maxerr(res[("lower", 1)], res[("lower", 0)]), "ak", P.maxerr(res[("ak", 1)], res[("ak", 0)]))
d = (res[("lower", 1)] - res[("lower", 0)]).abs()
print("lower diff per row max:", d.max(dim=1).values.reshape(b, 20))
print("lower diff per col max:", d.max(dim=0).values)
