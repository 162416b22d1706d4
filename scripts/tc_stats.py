import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mmego_b200 import _capi, synth
h = _capi.Handle(); h.set_weights(_capi.NET_IMU, synth.imu_state_dict(0))
imu = synth.batch(2048, seed=3)["imu"].cuda()
import itertools
for pair, chunk, extra in [(1, 4, 0), (1, 8, 0), (0, 4, 0)]:
    if True:
        h.set_option("tc_cta_pair", pair); h.set_option("tc_kb_chunk", chunk); h.set_option("tc_dbg", 0)
        h.imu_forward(imu); torch.cuda.synchronize()
        h.set_option("tc_dbg", 4 | extra); h.debug_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); h.imu_forward(imu); e1.record(); torch.cuda.synchronize()
        s = h.debug_stats()
        tiles = max(1, s[3])
        print(f"pair={pair} chunk={chunk}: {e0.elapsed_time(e1):.1f} ms | MMA thread per work item: total {s[6]/tiles:.0f} clk, "
              f"waiting epilogue {s[4]/tiles:.0f}, waiting TMA {s[5]/tiles:.0f}   (items {tiles})")
