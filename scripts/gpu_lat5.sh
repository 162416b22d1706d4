#!/bin/bash
# tagged-word exchange of the resident LSTM kernel: parity, then A/B against the fence + arrival-counter exchange on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "latency or graph_replay or smoke" 2>&1 | tail -4
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/lat5.log | tail -24
import sys, time, torch
sys.path.insert(0, ".")
from mmego_b200.Processor.Test.Demo_test import MMEgo
from tests import _parity as P
from oracle import mmego_oracle as O
h = P.make_handle()
for B in (1, 2, 3, 4):
    imu = O.synth_batch(B, seed=B)["imu"].cuda()
    row = []
    for xc in (0, 1):
        h.set_option("imu_res_xchg", xc)
        for _ in range(3): out = h.imu_forward(imu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): out = h.imu_forward(imu)
        e1.record(); torch.cuda.synchronize()
        row.append((e0.elapsed_time(e1) / 20, out))
    same = torch.equal(row[0][1][0], row[1][1][0]) and torch.equal(row[0][1][1], row[1][1][1])
    print(f"IMU_Net B={B}: counter exchange {row[0][0]:.3f} ms per call, tagged words {row[1][0]:.3f} ms per call; bit-identical: {same}; error flag {h.debug_stats(reset=False)[7]}")
h.close()
for rnd in (0, 1):
    for xc in (0, 1):
        for bs in (1, 3):
            m = MMEgo(batch_size=bs, imu_surrogate=False, quiet=True)
            m.pipe.handle.set_option("imu_res_xchg", xc)
            m.eval_model()
            best = 1e9
            for _ in range(2):
                m.eval_model(); best = min(best, m.seconds)
            n = m.data.shape[0]
            print(f"imu_res_xchg={xc} batch={bs} graphed={m.graphed}: {best / n * 1e3:.3f} ms per snippet, {n / best:.0f} it/s, mpjpe {m.report['mpjpe_cm']:.6f}")
PY
