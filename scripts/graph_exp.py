"""GPU experiment: CUDA-graph capture of the device-side pipeline call for small batches (the reference's own eval
loop is batch 1, ~120 launches per snippet): eager vs graph replay time, results compared bit for bit."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmego_b200 import synth
from mmego_b200.pipeline import MMEgoPipeline

dev = torch.device("cuda:0")
pipe = MMEgoPipeline(dev, imu_state=synth.imu_state_dict(0))
for B in (1, 4, 16):
    sb = synth.batch(B, seed=7)
    imu, data0, skl = sb["imu"].to(dev), sb["data"].to(dev), sb["skl"].to(dev)
    data = data0.clone()
    for _ in range(3):
        data.copy_(data0)
        ref = pipe.forward(imu, data, skl).clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        data.copy_(data0)
        pipe.forward(imu, data, skl)
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / 50
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        data.copy_(data0)
        pipe.forward(imu, data, skl)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        data.copy_(data0)
        out = pipe.forward(imu, data, skl)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    same = torch.equal(out, ref)
    t0 = time.perf_counter()
    for _ in range(50):
        g.replay()
    torch.cuda.synchronize()
    graph = (time.perf_counter() - t0) / 50
    print(f"B={B}: eager {eager * 1e3:.3f} ms, graph replay {graph * 1e3:.3f} ms per call, bit-identical {same}", flush=True)
