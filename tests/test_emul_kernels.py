"""CPU (`not gpu`): the SAME .cu sources that nvcc builds for sm_100a, compiled by g++ against the execution-model
emulator of tests/emul/ (fibers for threads, cooperative __syncthreads / shuffles), driven through the C ABI and
compared with the reference-generated goldens and the oracle.  This checks kernel logic, indexing, weight packing and
the host schedules before GPU minutes are spent; it is test infrastructure and is never loaded by the product path."""
import os
import sys

import pytest
import torch

from mmego_b200 import _capi

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_emul  # noqa: E402

from . import _parity as P  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return _capi.Lib(build_emul.build())


@pytest.fixture(scope="module")
def handle(lib):
    h = P.make_handle(lib=lib, require_cuda=False)
    # The small batches of this suite take IMU_Net's latency path.  Its tensor-core form is emulated MMA by MMA (a 32-lane
    # exchange each): it is switched on only where it is the subject (test_imu_latency_path...), everything else runs the
    # exact-fp32 form of the same kernel so that the CPU suite stays at a few minutes.
    h.set_option("imu_res_tc", 0)
    yield h
    h.close()


def test_upper_lower_synth3(handle):
    P.check_upper_lower_golden(handle, "synth3.npz")


def test_upper_lower_real_sample(handle):
    P.check_upper_lower_golden(handle, "sample16.npz", sl=[slice(0, 1), slice(7, 8)])


@pytest.mark.parametrize("name", P.GCN_GOLDENS)
def test_gcn(handle, name):
    P.check_gcn_golden(handle, name=name)


def test_upper_lower_sweep_shape(handle):
    # (L, N) = (20, 512): 4x the config's point count, against the reference's own classes
    P.check_sweep_golden(handle, "sweep_L20_N512.npz")


@pytest.mark.parametrize("N", [64, 100, 128, 257, 300, 512])
def test_top64_bitonic_sort_path_matches_rank_select(handle, N):
    P.check_top64_sort_path(handle, handle, N)


def test_nan_inputs_stay_in_their_snippet(handle):
    P.check_nan_inputs_do_not_corrupt(handle)


def test_transforms(handle):
    P.check_transforms(handle)


def test_metrics(handle):
    P.check_metrics(handle)


def test_imu_golden(handle):
    P.check_imu_golden(handle, "synth", 1)


def test_pipeline_ragged_shapes(handle):
    # L, N, n_imu away from the config values; B*L not a multiple of any tile; distinct skeletons (F8)
    P.check_pipeline_vs_oracle(handle, B=3, L=5, N=70, n_imu=3, seed=21)
    P.check_pipeline_vs_oracle(handle, B=3, L=5, N=70, n_imu=3, seed=21, truth64=True)      # judged against float64


def test_pipeline_many_snippets(handle):
    # enough sequences that the H=64 recurrence runs with 4 tiles per CTA (the emulated device has 4 SMs), ragged tail
    P.check_pipeline_vs_oracle(handle, B=136, L=2, N=64, n_imu=1, seed=5)


def test_snippet_builder(handle):
    P.check_snippet_builder(handle)


def test_posepc_windows_split(handle, tmp_path):
    """Host logic of the PosePC mirror: windows from the end of each recording, seeded shuffle, 80/20 split."""
    import numpy as np
    from mmego_b200.Util.Universal_Util.Dataset_sample import PosePC, snippet_windows
    from oracle import mmego_oracle as O
    z = dict(np.load(os.path.join(P.GOLDEN, "raw_subset.npz")))
    assert np.array_equal(snippet_windows(z["rec_start"], 20), O.snippet_windows(z["rec_start"]))
    path = os.path.join(tmp_path, "raw.npz")
    np.savez(path, **{k: z[k] for k in ("points", "pt_start", "key", "imu", "R_btc", "t_R0R", "R_ref", "orientation_ref",
                                        "rec_start", "skl")})
    vis = PosePC(train=False, vis=True, packed_path=path, lib_handle=handle)
    tr = PosePC(train=True, vis=False, packed_path=path, lib_handle=handle)
    te = PosePC(train=False, vis=False, packed_path=path, lib_handle=handle)
    assert len(vis) == 12 and len(tr) == 9 and len(te) == 3
    assert sorted(np.concatenate([tr.starts, te.starts]).tolist()) == sorted(vis.starts.tolist())
    perm = np.arange(12)
    np.random.RandomState(1).shuffle(perm)                     # Config.dataset_random_seed of the reference
    assert np.array_equal(np.concatenate([tr.starts, te.starts]), vis.starts[perm])
    b = vis.batch([0, 5])
    assert b["data"].shape == (2, 20, 128, 6) and b["skl"].shape == (2, 20, 3)
    item = vis[5]
    assert len(item) == 9 and len(te[0]) == 8                 # the reference's tuple lengths (vis / test)
    assert np.array_equal(item[0], b["data"][1].cpu().numpy())
    assert np.array_equal(item[1], z["exp_key"][5])
    assert np.array_equal(item[6], z["exp_R"][5]) and item[7].shape == (20, 1, 3)
    # the loader fields the networks never read, against the reference's own PosePC (tests/golden/extras_subset.npz)
    ex = dict(np.load(os.path.join(P.GOLDEN, "extras_subset.npz")))
    ex_path = os.path.join(tmp_path, "extras.npz")
    np.savez(ex_path, ground=ex["ground"], foot_contact_raw=ex["foot_contact_raw"])
    vis2 = PosePC(train=False, vis=True, packed_path=path, lib_handle=handle, extras_path=ex_path)
    for i in (0, 5, 11):
        it = vis2[i]
        assert np.array_equal(it[4], ex["exp_ground"][i]) and np.array_equal(it[5], ex["exp_foot_contact"][i])
        assert np.abs(it[8] - ex["exp_R_RtW"][i]).max() < 1e-12


def test_errors(handle):
    P.check_errors(handle)


def test_sharded_body_index(handle):
    """A shard (b_offset, B_global) reproduces rows of the unsharded call, including initial_body[r % B]."""
    g = P.golden("synth3.npz")
    h0 = torch.zeros(6, 1, 64)
    for b in range(3):
        x = g["data"][b:b + 1].clone()
        l = handle.upper_forward(x, h0, h0.clone(), g["skl"], g["R"][b:b + 1].contiguous(), g["t"][b:b + 1].contiguous(),
                                 b_offset=b, B_global=3)[0]
        assert P.maxerr(l, g["upper_l"][b:b + 1]) < P.POS_TOL
        ll = handle.lower_forward(g["upper_l"][b:b + 1].contiguous(), g["x1"][b:b + 1].clone(), g["skl"],
                                  g["R"][b:b + 1].contiguous(), g["t"][b:b + 1].contiguous(), b_offset=b, B_global=3)[0]
        assert P.maxerr(ll, g["lower_l"][b:b + 1]) < P.POS_TOL


def test_infer_host_chunk_pipeline_matches_device_path(handle):
    """mmego_infer_host cuts the batch into host_chunk-sized stages (b_offset/B_global keep the reference's
    initial_body[r % B] indexing of the whole batch); results equal the one-shot device call."""
    from oracle import mmego_oracle as O
    sb = O.synth_batch(3, L=4, N=70, n_imu=2, seed=9, distinct_skeletons=True)
    pred_d = handle.pipeline_forward(sb["imu"], sb["data"].clone(), sb["skl"])
    tg = (pred_d + 0.01).contiguous()
    handle.set_option("host_chunk", 2)
    try:
        pred_h, sums = handle.infer_host(sb["imu"], sb["data"], sb["skl"], tg)
    finally:
        handle.set_option("host_chunk", 2048)
    assert torch.equal(pred_h, pred_d)
    assert sums[43].item() == 12
    # ramp-up: a short first stage (host_chunk / 8), then full stages, then the tail
    sb = O.synth_batch(11, L=2, N=64, n_imu=1, seed=4, distinct_skeletons=True)
    pred_d = handle.pipeline_forward(sb["imu"], sb["data"].clone(), sb["skl"])
    handle.set_option("host_chunk", 8)
    try:
        pred_h, sums = handle.infer_host(sb["imu"], sb["data"], sb["skl"], (pred_d + 0.01).contiguous())
    finally:
        handle.set_option("host_chunk", 2048)
    assert torch.equal(pred_h, pred_d)
    assert sums[43].item() == 22


def test_boundary_validates_sizes_before_the_library_reads_them(handle):
    """The C entry points trust their sizes; the ctypes layer checks everything they will read."""
    from oracle import mmego_oracle as O
    sb = O.synth_batch(2, L=3, N=64, n_imu=2, seed=1)
    with pytest.raises(_capi.MMEgoError, match="initial_body"):        # a shard passing its LOCAL skeletons with the global B
        handle.pipeline_forward(sb["imu"], sb["data"].clone(), sb["skl"], b_offset=2, B_global=4)
    with pytest.raises(_capi.MMEgoError, match="imu must be"):
        handle.pipeline_forward(sb["imu"][:1].contiguous(), sb["data"].clone(), sb["skl"])
    with pytest.raises(_capi.MMEgoError, match="target must be"):
        handle.pipeline_forward(sb["imu"], sb["data"].clone(), sb["skl"], torch.zeros(2, 3, 20, 3), torch.zeros(46, dtype=torch.float64))
    with pytest.raises(_capi.MMEgoError, match="initial_body"):
        handle.infer_host(sb["imu"], sb["data"], sb["skl"][:1].contiguous())
    import numpy as np
    z = dict(np.load(os.path.join(P.GOLDEN, "raw_subset.npz")))
    raw = {k: torch.from_numpy(np.ascontiguousarray(z[k])) for k in _capi.RawFramesStruct.DTYPES}
    n_frames = raw["pt_start"].numel() - 1
    with pytest.raises(_capi.MMEgoError, match="raw frames"):
        handle.build_snippets(raw, torch.tensor([n_frames - 5], dtype=torch.int64))


def test_imu_latency_path_matches_oracle_and_ffma_path(handle):
    """The resident-weights fp32 LSTM kernels (small-batch latency path) on the emulator -- one launch per timestep there,
    the same kernel code -- against the oracle and against the fp32 FFMA generation."""
    from oracle import mmego_oracle as O
    # (1, 23, 2): rnn_slow's up-front input pass has a ragged second block.  tc: rnn_fast on (emulated) mma.sync fragments --
    # two m-tiles at 20 sequences, two blocks of sequences at 22 (20 + 2), one m-tile at 15 -- kept to few timesteps (see `handle`)
    for B, L, n, tc in ((1, 20, 20, 0), (1, 20, 2, 1), (3, 5, 3, 1), (2, 11, 2, 1), (3, 20, 2, 0), (1, 23, 2, 0)):
        sb = O.synth_batch(B, L=L, N=64, n_imu=n, seed=5 + B)
        handle.set_option("imu_res_tc", tc)
        try:
            R1, t1 = handle.imu_forward(sb["imu"])
        finally:
            handle.set_option("imu_res_tc", 0)
        Rr, tr = O.imu_forward(O.synth_imu_state_dict(0), sb["imu"])
        assert P.rot_angle_deg(R1, Rr) < P.ANG_TOL / 2 and P.maxerr(t1, tr) < 1e-6
        if (B, L, n) in ((3, 5, 3), (3, 20, 2)):        # ... and against the fp32 FFMA generation (the emulated big-LSTM path is slow)
            handle.set_option("imu_resident", 0)
            try:
                R0, t0 = handle.imu_forward(sb["imu"])
            finally:
                handle.set_option("imu_resident", 1)
            assert P.rot_angle_deg(R1, R0) < P.ANG_TOL / 2


def test_imu_latency_path_option_forms_agree(handle):
    """Every selectable form of the resident kernel is the same math.  Promised: (a) tagged words vs arrival counter as the
    exchange of h: bit-identical; (b) staged vs direct tensor-core operands, up-front vs in-step input projections, and the
    exact-fp32 form: different summation orders / operand splits of the same products, far inside the tolerance."""
    from oracle import mmego_oracle as O
    sb = O.synth_batch(1, L=3, N=64, n_imu=2, seed=21)
    Rr, tr = O.imu_forward(O.synth_imu_state_dict(0), sb["imu"])
    base = dict(imu_res_tc=1, imu_res_direct=1, imu_res_xchg=1, imu_res_pre=1)
    out = {}
    try:
        for name, kw in (("default", {}), ("counter", dict(imu_res_xchg=0)), ("staged", dict(imu_res_direct=0)),
                         ("staged_counter", dict(imu_res_direct=0, imu_res_xchg=0)), ("in_step", dict(imu_res_pre=0)),
                         ("fp32", dict(imu_res_tc=0))):
            for k, v in {**base, **kw}.items():
                handle.set_option(k, v)
            out[name] = handle.imu_forward(sb["imu"])
    finally:
        for k, v in {**base, "imu_res_tc": 0}.items():          # the fixture's setting
            handle.set_option(k, v)
    for a, b in (("default", "counter"), ("staged", "staged_counter")):
        assert torch.equal(out[a][0], out[b][0]) and torch.equal(out[a][1], out[b][1]), (a, b)
    for name, (R, t) in out.items():
        assert P.rot_angle_deg(R, Rr) < P.ANG_TOL / 2 and P.maxerr(t, tr) < 1e-6, name
