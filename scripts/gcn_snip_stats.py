"""MMA-thread cycle accounting of tconv_snip_kernel (test-only variant library): per layer, where the issuing thread waits."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
from mmego_b200 import _capi, build as B_
h = P.make_handle(lib=_capi.Lib(B_.VARIANTS["ffma"]["lib"]), with_imu=False)
B = 2048
for kv in sys.argv[1:]:
    k, v = kv.split("="); h.set_option(k, int(v))
torch.manual_seed(0)
x = (torch.randn(B, 3, 20, 15, 1, device="cuda") * 0.5).contiguous()
for _ in range(2):
    h.gcn_extract_feature(x)
h.debug_stats(reset=True)
h.gcn_extract_feature(x)
st = h.debug_stats(reset=True)
n = max(1, st[4])
print(f"MMA threads: {n} (3 launches x grid); avg cycles per launch-thread: total {st[0]/n:.0f}, wait drain {st[1]/n:.0f}, wait window {st[2]/n:.0f}, wait weights {st[3]/n:.0f}; err {st[7]}")
