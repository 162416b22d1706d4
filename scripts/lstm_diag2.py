"""Dump (emulator) or compare (GPU) the per-layer gx / y of the lower small-LSTM stack in mma mode."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmego_b200 import _capi
from tests import _parity as P
gpu = "--gpu" in sys.argv
if gpu:
    h = P.make_handle(with_imu=False)
else:
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul
    h = P.make_handle(lib=_capi.Lib(build_emul.build()), require_cuda=False, with_imu=False)
g = P.golden("synth3.npz")
b = 3
h.set_option("gcn_gemm", 0)
x = P.dev(h, g["x1"].clone())
skl, R, t = P.dev(h, g["skl"]), P.dev(h, g["R"]), P.dev(h, g["t"])
taps = {}
for l in range(3):
    taps["small_lstm.gx%d" % l] = torch.zeros(b * 20, 512, device=h.device)
    taps["small_lstm.y%d" % l] = torch.zeros(b * 20, 128, device=h.device)
taps["lower.ak"] = torch.zeros(b * 20, 192, device=h.device)
for k, v in taps.items():
    h.tap(k, v)
ll = h.lower_forward(P.dev(h, g["upper_l"]), x, skl, R, t)[0]
if gpu:
    torch.cuda.synchronize()
out = {k: v.cpu() for k, v in taps.items()}
path = os.path.join(ROOT, "scripts", "lstm_diag2_emul.pt")
if not gpu:
    torch.save(out, path)
    print("saved", path, "lower_l vs golden", P.maxerr(ll, g["lower_l"]))
else:
    ref = torch.load(path)
    for k in out:
        d = (out[k] - ref[k]).abs()
        print(k, "max|d| %.3e" % float(d.max()), "max|ref| %.3e" % float(ref[k].abs().max()))
        if d.max() > 1e-3:
            bad = (d > 1e-3)
            print("   bad rows:", bad.any(dim=1).nonzero().flatten().tolist()[:40])
            print("   bad cols:", bad.any(dim=0).nonzero().flatten().tolist()[:64])
if gpu:
    gx = out["small_lstm.gx0"].reshape(b, 20, 2, 8, 4, 8)      # seq, t, dir, u, gate, n
    y = out["small_lstm.y0"].reshape(b, 20, 2, 8, 8)
    yr = ref["small_lstm.y0"].reshape(b, 20, 2, 8, 8)
    pre = gx[:, 0, 0]                                           # forward dir, first step: h0 = c0 = 0
    i, f, gg, o = pre[:, :, 0], pre[:, :, 1], pre[:, :, 2], pre[:, :, 3]
    c = torch.sigmoid(i) * torch.tanh(gg)
    hexp = torch.sigmoid(o) * torch.tanh(c)
    print("step0 fwd: gpu vs expected-from-gx %.3e ; emul vs expected %.3e" % (float((y[:, 0, 0] - hexp).abs().max()), float((yr[:, 0, 0] - hexp).abs().max())))
    print("seq0 u0 pre i", i[0, 0], "\n g", gg[0, 0], "\n o", o[0, 0])
    print("gpu  h", y[0, 0, 0, 0], "\nemul h", yr[0, 0, 0, 0], "\nexp  h", hexp[0, 0])
    for tt in range(0, 20, 4):
        print("t", tt, "fwd diff %.3e" % float((y[:, tt, 0] - yr[:, tt, 0]).abs().max()), "bwd diff %.3e" % float((y[:, tt, 1] - yr[:, tt, 1]).abs().max()))
if gpu:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    torch.save(out, os.path.join(ROOT, "gpurun_out", "lstm_taps_gpu.pt"))
