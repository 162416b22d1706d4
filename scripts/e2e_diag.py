"""Where does the e2e wall clock go?  Times mmego_infer_host (C call only) against the Python wrapper and the device path."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmego_b200 import synth, _capi
from mmego_b200.pipeline import MMEgoPipeline, SUMS_LEN
import ctypes as C
dev = torch.device("cuda", 0)
pipe = MMEgoPipeline(dev, imu_state=None)
B = 4096
sb = synth.batch(B, seed=1234)
imu_h, data_h, skl_h = sb["imu"].pin_memory(), sb["data"].pin_memory(), sb["skl"].pin_memory()
pred0 = pipe.forward(imu_h.to(dev), data_h.to(dev), skl_h.to(dev))
tg = synth.target_like(pred0, seed=99).contiguous().pin_memory()
h = pipe.handle
for hc in (2048, 4096, 1024):
    h.set_option("host_chunk", hc)
    for _ in range(2):
        pipe.infer_host(imu_h, data_h, skl_h, tg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        pipe.infer_host(imu_h, data_h, skl_h, tg)
    t_py = (time.perf_counter() - t0) / 4
    pred = torch.empty(B, 20, 21, 3, dtype=torch.float32, pin_memory=True)
    sums = torch.zeros(SUMS_LEN, dtype=torch.float64, pin_memory=True)
    t0 = time.perf_counter()
    for _ in range(4):
        rc = h.lib.dll.mmego_infer_host(h._h, imu_h.data_ptr(), data_h.data_ptr(), skl_h.data_ptr(), tg.data_ptr(), pred.data_ptr(),
                                        sums.data_ptr(), B, 20, 128, 20, 0, 0, B)
    t_c = (time.perf_counter() - t0) / 4
    print(f"host_chunk {hc}: python wrapper {t_py*1e3:.2f} ms, C call only {t_c*1e3:.2f} ms")
# device path for comparison (same process / clocks)
imu_d, data0, skl_d, tg_d = imu_h.to(dev), data_h.to(dev), skl_h.to(dev), tg.to(dev)
data_d = torch.empty_like(data0)
sums_d = torch.zeros(SUMS_LEN, dtype=torch.float64, device=dev)
for _ in range(2):
    data_d.copy_(data0); pipe.forward(imu_d, data_d, skl_d, tg_d, sums_d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    data_d.copy_(data0); pipe.forward(imu_d, data_d, skl_d, tg_d, sums_d)
torch.cuda.synchronize()
print(f"device path: {(time.perf_counter()-t0)/4*1e3:.2f} ms")
