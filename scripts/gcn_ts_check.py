"""GPU check of the tensor-memory-weights variant of the snippet-resident temporal conv (gcn_snip = 9) against the
reference-generated ST-GCN vectors, and its time against the shared-memory-weights variant (gcn_snip = 1)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _parity as P
h = P.make_handle(with_imu=False)
torch.manual_seed(0)
x = (torch.randn(2048, 3, 20, 15, 1, device="cuda") * 0.5).contiguous()
outs = {}
for v in (1, 9):
    h.set_option("gcn_snip", v)
    errs = []
    for name in P.GCN_GOLDENS:
        try:
            errs.append(P.check_gcn_golden(h, name=name))
        except AssertionError as e:
            errs.append("FAIL " + str(e)[:60])
    for _ in range(2):
        outs[v] = h.gcn_extract_feature(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        outs[v] = h.gcn_extract_feature(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"gcn_snip={v}: goldens {errs}; {e0.elapsed_time(e1) / 5:.3f} ms per 2048 snippets; error flag {h.debug_stats(reset=True)[7]}")
print("max |snip9 - snip1| / max:", float((outs[9] - outs[1]).abs().max() / outs[1].abs().max()))
