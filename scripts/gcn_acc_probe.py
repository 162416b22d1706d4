"""Which stage sets the lower-joint error at large sample counts?  B = 200 synthetic snippets against the float64 oracle
with the temporal-conv variants (gcn_snip = 1 default, 17 = second drain group per block, 0 = row-tiled GEMM) and with the
IMU head pose taken from the oracle (isolates Upper/Lower from IMU_Net's rotation error)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import _parity as P
O = P.O
B = 200
sb = O.synth_batch(B, seed=12, distinct_skeletons=True)
up_sd, lo_sd = P.checkpoints()
imu_sd = O.synth_imu_state_dict(0)
ref64 = O.pipeline(imu_sd, up_sd, lo_sd, sb["imu"], sb["data"], sb["skl"], dtype=torch.float64)
ref32 = O.pipeline(imu_sd, up_sd, lo_sd, sb["imu"], sb["data"], sb["skl"])
h = P.make_handle()
out = {}
VARIANTS = [int(a) for a in sys.argv[1:]] or [1, 17, 0]
xg = (torch.randn(2048, 3, 20, 15, 1, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0)) * 0.5).contiguous()
for snip in VARIANTS:
    h.set_option("gcn_snip", snip)
    outs = dict(R=torch.empty(B, 20, 3, 3, device=h.device), t=torch.empty(B, 20, 3, device=h.device),
                upper_l=torch.empty(B, 20, 15, 3, device=h.device), lower_l=torch.empty(B, 20, 8, 3, device=h.device))
    h.pipeline_forward(P.dev(h, sb["imu"]), P.dev(h, sb["data"].clone()), P.dev(h, sb["skl"]), outs=outs)
    # Upper/Lower alone with the float64 oracle's head pose (fp32-rounded)
    R64, t64 = ref64["R"].float(), ref64["t"].float()
    x = P.dev(h, sb["data"].clone())
    h0 = torch.zeros(6, B, 64, device=h.device)
    up = h.upper_forward(x, h0, h0.clone(), P.dev(h, sb["skl"]), P.dev(h, R64), P.dev(h, t64))[0]
    lo = h.lower_forward(up, x, P.dev(h, sb["skl"]), P.dev(h, R64), P.dev(h, t64))[0]
    ref64b = O.pipeline(None, up_sd, lo_sd, None, sb["data"], sb["skl"], R_t=(R64.double(), t64.double()), dtype=torch.float64)
    out[snip] = dict(lower64=P.maxerr(outs["lower_l"].double(), ref64["lower_l"]), upper64=P.maxerr(outs["upper_l"].double(), ref64["upper_l"]),
                     lower32=P.maxerr(outs["lower_l"], ref32["lower_l"]),
                     lower64_given_R=P.maxerr(lo.double(), ref64b["lower_l"]), upper64_given_R=P.maxerr(up.double(), ref64b["upper_l"]))
    for _ in range(2):
        h.gcn_extract_feature(xg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        h.gcn_extract_feature(xg)
    e1.record()
    torch.cuda.synchronize()
    out[snip]["gcn_ms_per_2048"] = e0.elapsed_time(e1) / 5
    print(snip, {k: (round(v, 9) if v < 1 else round(v, 3)) for k, v in out[snip].items()}, flush=True)
h.set_option("gcn_snip", 1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({str(k): v for k, v in out.items()}, open(os.path.join(ROOT, "gpurun_out", "gcn_acc_probe.json"), "w"), indent=1)
