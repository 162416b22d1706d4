"""CPU (`not gpu`): host-side logic -- checkpoint-layout contract of the drop-in modules, synthetic generators,
report arithmetic, and the N>1 data-parallel path (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmego_b200 import _capi, synth
from mmego_b200.engine import MMEgoError
from mmego_b200.Net.GCN import Model as GcnModel
from mmego_b200.Net.IMU_Net import IMUNet
from mmego_b200.Net.Lower_Net import LowerNet
from mmego_b200.Net.Upper_Net import UpperNet
from mmego_b200.pipeline import ShardedRunner, report_from_sums, shard_bounds
from oracle import mmego_oracle as O


def test_state_dict_layout_matches_shipped_checkpoints(checkpoints):
    up, lo = checkpoints
    for net, sd in ((UpperNet(), up), (LowerNet(64), lo)):
        mine = net.state_dict()
        assert list(mine.keys()) == list(sd.keys())
        for k, v in sd.items():
            assert mine[k].shape == v.shape and mine[k].dtype == v.dtype, k
        net.load_state_dict(sd, strict=True)
        assert all(torch.equal(net.state_dict()[k], v) for k, v in sd.items())


def test_imu_layout_and_roundtrip(tmp_path):
    net = IMUNet(15, 9, 512, 2, True, 0.1)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == O.imu_state_dict_shapes()
    assert sum(v.numel() for v in net.state_dict().values()) == 23119912
    p = net.save(str(tmp_path / "imu.pth"))
    other = IMUNet(15, 9, 512, 2, True, 0.1)
    other.load(p)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), other.state_dict().values()))


def test_gcn_adjacency_buffer(checkpoints):
    _, lo = checkpoints
    g = GcnModel(3, 64, {"layout": "kinect_upper", "strategy": "distance"})
    assert float((g.state_dict()["A"] - lo["keyEncoder.gcn.A"]).abs().max()) < 1e-7
    assert [k for k in g.state_dict()] == [k[len("keyEncoder.gcn."):] for k in lo if k.startswith("keyEncoder.gcn.")]


def test_unsupported_configs_and_cpu_inputs_raise():
    with pytest.raises(MMEgoError):
        IMUNet(15, 9, 256, 2, True, 0.1)
    with pytest.raises(MMEgoError):
        LowerNet(32)
    with pytest.raises(MMEgoError, match="no CPU path"):
        UpperNet()(torch.zeros(1, 20, 128, 6), None, None, None, None, None)
    with pytest.raises(MMEgoError, match="no CPU path"):
        IMUNet(15, 9, 512, 2, True, 0.1)(torch.zeros(1, 20, 20, 15))


def test_synth_matches_oracle_generator():
    a, b = synth.batch(2, seed=5, distinct_skeletons=True), O.synth_batch(2, seed=5, distinct_skeletons=True)
    assert all(torch.equal(a[k], b[k]) for k in a)
    x, y = synth.imu_state_dict(3), O.synth_imu_state_dict(3)
    assert list(x) == list(y) and all(torch.equal(x[k], y[k]) for k in x)
    frac_zero = float((a["data"].abs().sum(-1) == 0).float().mean())
    assert 0.25 < frac_zero < 0.55


def test_report_from_sums_matches_oracle(golden_dir):
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "sample16.npz")).items()}
    s = O.metric_sums(g["pred"], g["upper_l"], g["lower_l"], g["target"])
    flat = np.zeros(_capi.SUMS_LEN)
    flat[0:21], flat[21], flat[22], flat[23:43], flat[43] = s["err_joint"], s["err_upper"], s["err_lower"], s["angle_bone"], s["frames"]
    rep, want = report_from_sums(flat), O.report_from_sums(s)
    for k in ("mpjpe_cm", "upper_cm", "lower_cm", "angle_deg"):
        assert abs(rep[k] - want[k]) < 1e-12
    assert np.allclose(rep["per_joint_cm"], want["per_joint_cm"])


def test_shard_bounds_cover_everything():
    for B in (1, 7, 8, 4096, 835):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(B, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(B * 2 * 21 * 3, dtype=torch.float32).reshape(B, 2, 21, 3)

    def step(lo, hi, Bg):      # stand-in for the CUDA pipeline: a deterministic function of the global snippet index
        sums = torch.zeros(_capi.SUMS_LEN, dtype=torch.float64)
        sums[43] = (hi - lo) * 2
        sums[0] = float(full[lo:hi].sum())
        return full[lo:hi].clone(), sums

    runner = ShardedRunner(step, world, rank)
    pred, sums = runner.run(B)
    ok = torch.equal(pred, full) and sums[43].item() == B * 2 and abs(sums[0].item() - float(full.sum())) < 1e-3
    # pipelined form: two steps outstanding, collected in order, results identical to the synchronous form
    runner.submit(B)
    runner.submit(B)
    for _ in range(2):
        p2, s2 = runner.collect()
        ok = ok and torch.equal(p2, full) and s2[43].item() == B * 2
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_sharded_runner_gloo_world2(B):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    assert out[0] and out[1]
