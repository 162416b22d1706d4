"""CPU: the oracle restatement (oracle/mmego_oracle.py) against vectors frozen from the reference's own
classes (oracle/make_golden.py).  Tolerances: the two are different fp32 evaluation orders of the same
math, so agreement is at fp32 noise (4e-7 m positions, 3e-6 on R entries)."""
import os

import numpy as np
import pytest
import torch

from oracle import mmego_oracle as O


def _load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _maxerr(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def test_graph_adjacency_matches_checkpoint_buffer(checkpoints):
    _, lo = checkpoints
    assert _maxerr(torch.from_numpy(O.graph_adjacency()).float(), lo["keyEncoder.gcn.A"]) < 1e-7


@pytest.mark.parametrize("name", ["gcn2.npz", "gcn_T40.npz", "gcn_T80.npz"])
def test_gcn_extract_feature(golden_dir, checkpoints, name):
    """T = 20 (config) and T = 40 / 80 (sweep): the reference's GCN.Model accepts any T (Net/GCN.py:103-117)."""
    _, lo = checkpoints
    g = _load(golden_dir, name)
    sd = {k[len("keyEncoder.gcn."):]: v for k, v in lo.items() if k.startswith("keyEncoder.gcn.")}
    out = O.gcn_extract_feature(sd, g["x"])
    assert out.shape == g["out"].shape
    assert _maxerr(out, g["out"]) < 2e-5 * float(g["out"].abs().max())


@pytest.mark.parametrize("name", ["synth3.npz", "sample16.npz"])
def test_upper_lower_against_reference(golden_dir, checkpoints, name):
    up_sd, lo_sd = checkpoints
    g = _load(golden_dir, name)
    B = g["data"].shape[0]
    one_call = name == "synth3.npz"          # synth3 was produced by ONE reference call with B=3 (F8 visible)
    chunks = [slice(0, B)] if one_call else [slice(i, i + 1) for i in range(B)]
    for sl in chunks:
        b = g["data"][sl].shape[0]
        h0 = torch.zeros(6, b, 64)
        l, q, w, hn, cn, x1 = O.upper_forward(up_sd, g["data"][sl], h0, h0, g["skl"][sl], g["R"][sl], g["t"][sl])
        assert _maxerr(l, g["upper_l"][sl]) < 5e-6
        assert _maxerr(q, g["q_upper"][sl]) < 2e-5
        assert _maxerr(w.reshape(b, 20, -1), g["gw"][sl]) < 1e-5
        assert _maxerr(hn.permute(1, 0, 2), g["hn"][sl]) < 1e-5
        assert _maxerr(cn.permute(1, 0, 2), g["cn"][sl]) < 1e-5
        assert _maxerr(x1, g["x1"][sl]) < 2e-6
        # feed the reference's own upper_l / mutated cloud to isolate the lower stage
        # tie_rule="torch_cpu" replays the CPU sort backend the golden was produced with: the sample data has
        # distinct points with identical xyz, so the unstable sort's tie choice is visible (up to ~2 mm).
        lo, ql, x2 = O.lower_forward(lo_sd, g["upper_l"][sl], g["x1"][sl], g["skl"][sl], g["R"][sl], g["t"][sl],
                                     tie_rule="torch_cpu")
        assert _maxerr(x2, g["x2"][sl]) < 5e-6
        assert _maxerr(lo, g["lower_l"][sl]) < 1e-5
        assert _maxerr(ql, g["q_lower"][sl]) < 5e-5
        # the framework's documented rule (lowest slot wins) stays within the tie-induced spread
        lo_s = O.lower_forward(lo_sd, g["upper_l"][sl], g["x1"][sl], g["skl"][sl], g["R"][sl], g["t"][sl])[0]
        assert _maxerr(lo_s, g["lower_l"][sl]) < (1e-5 if one_call else 5e-3)
        pred = O.assemble(l, lo)
        assert _maxerr(pred, g["pred"][sl]) < 1e-5


@pytest.mark.parametrize("name", ["sweep_L40_N256.npz", "sweep_L80_N128.npz", "sweep_L20_N512.npz"])
def test_upper_lower_sweep_shapes_against_reference(golden_dir, checkpoints, name):
    """Non-config (L, N): the oracle is pinned to the reference's own classes there too (one reference call per shape)."""
    up_sd, lo_sd = checkpoints
    g = _load(golden_dir, name)
    B = g["data"].shape[0]
    h0 = torch.zeros(6, B, 64)
    l, q, w, hn, cn, x1 = O.upper_forward(up_sd, g["data"], h0, h0, g["skl"], g["R"], g["t"])
    assert _maxerr(l, g["upper_l"]) < 5e-6
    assert _maxerr(q, g["q_upper"]) < 2e-5
    assert _maxerr(w.reshape(B, g["data"].shape[1], -1), g["gw"]) < 1e-5
    assert _maxerr(hn.permute(1, 0, 2), g["hn"]) < 1e-5
    lo, ql, _ = O.lower_forward(lo_sd, g["upper_l"], x1, g["skl"], g["R"], g["t"], tie_rule="torch_cpu")
    assert _maxerr(lo, g["lower_l"]) < 1e-5
    assert _maxerr(ql, g["q_lower"]) < 5e-5
    assert _maxerr(O.assemble(l, lo), g["pred"]) < 1e-5


def test_body_index_quirk_is_visible(golden_dir, checkpoints):
    """F8: with distinct skeletons the reference uses initial_body[r % B]; r // L must NOT match."""
    up_sd, _ = checkpoints
    g = _load(golden_dir, "synth3.npz")
    h0 = torch.zeros(6, 3, 64)
    l = O.upper_forward(up_sd, g["data"], h0, h0, g["skl"], g["R"], g["t"], ref_body_index=False)[0]
    assert _maxerr(l, g["upper_l"]) > 1e-3


def test_imu_against_reference(golden_dir):
    g = _load(golden_dir, "imu_seed0.npz")
    sd = O.synth_imu_state_dict(0)
    for tag in ("synth", "real"):
        R, t = O.imu_forward(sd, g["imu_" + tag])
        assert _maxerr(R, g["R_" + tag]) < 2e-5
        assert _maxerr(t, g["t_" + tag]) < 2e-5


def test_metrics_pin_on_slice(golden_dir):
    """metric_sums/report_from_sums reproduce the reference's mean-of-batch-means on equal-size batches."""
    g = _load(golden_dir, "sample16.npz")
    s = O.metric_sums(g["pred"], g["upper_l"], g["lower_l"], g["target"])
    rep = O.report_from_sums(s)
    e = (g["pred"] - g["target"]).square().sum(-1).sqrt()
    assert abs(rep["mpjpe_cm"] - float(e.mean()) * 100) < 1e-4
    assert np.allclose(rep["per_joint_cm"], e.mean((0, 1)).numpy() * 100, atol=1e-4)


def test_snippet_builder_oracle_matches_reference_loader():
    """oracle.build_snippets (CPU restatement of Dataset_sample.py:153-260) against tensors produced by the reference's
    own PosePC class on the first three recordings (scripts/pack_sample_data.py --fixture): bit-exact."""
    import numpy as np
    z = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raw_subset.npz")))
    starts = O.snippet_windows(z["rec_start"])
    assert len(starts) == len(z["exp_data"]) == 12
    slots = O.recover_slots(z, starts, z["exp_data"])
    out = O.build_snippets(z, starts, slots)
    for a, b in (("data", "exp_data"), ("imu", "exp_imu"), ("key", "exp_key"), ("R", "exp_R"), ("t", "exp_t"), ("skl", "exp_skl")):
        assert np.array_equal(out[a], z[b]), a
