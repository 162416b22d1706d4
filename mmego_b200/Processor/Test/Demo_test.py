"""Drop-in for the evaluation driver of the reference's Processor/Test/Demo_test.py:22-184 (class MMEgo).

Same constructor-less surface (`MMEgo().eval_model()`), same five printed lines (Demo_test.py:176-180) and the same
return tuple (:184).  Differences, all on the host side:
  * batches of `Config.batch_size` snippets instead of the hard-coded 1 (Demo_test.py:61); all batches are equal-sized
    or the tail is handled by exact sums, so the printed means are the global means the reference computes.  The
    reference's loop feeds ONE snippet per call, so its ForKinematics' `initial_body[r % B]` (B = 1) is that snippet's
    own skeleton: the driver therefore runs the networks in per-snippet body-index mode (`initial_body[r // L]`), which
    is identical to the reference at batch 1 for any batch size here -- also on multi-subject data, where replaying
    `r % B` with B > 1 would hand frames the skeletons of other snippets,
  * error sums are accumulated on the device and read back once, instead of six .item() syncs per batch,
  * the sample set is read from the frozen tensor file built with the reference loader (np.random.seed(0)), because
    the loader's pad-slot placement uses the unseeded global RNG (Dataset_sample.py:215-223); with `from_raw=True`
    (`main.py --infer --from_raw`) the batches are instead built on the GPU from the packed raw sensor cache by
    mmego_build_snippets (Util/Universal_Util/Dataset_sample.py of this package), seeded placement.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ...Config.config import Config
from ...engine import MMEgoError
from ...pipeline import SUMS_LEN, GraphedStep, MMEgoPipeline, report_from_sums


class MMEgo:
    GRAPH_MAX_SEQ = 120         # batch_size * frame_no up to which a step is replayed as a CUDA graph (the library's latency regime)

    def __init__(self, batch_size=None, device=None, imu_surrogate=None, quiet=False, from_raw=False, use_graph=True,
                 fused=True):
        self.device = torch.device(device or Config.device)
        if self.device.type != "cuda":
            raise MMEgoError("MMEgo needs a CUDA device (B200); there is no CPU fallback")
        self.batch_size = int(batch_size or Config.batch_size)
        self.frame_no = Config.frame_no
        self.pipe = MMEgoPipeline(self.device, body_index_mode="per_snippet")
        # With the IMU_Net checkpoint absent (it is missing from the reference mount) the only way to reproduce an
        # accuracy figure is the surrogate of SURVEY.md section 8(c): IMU_Net's own training targets as (R, t).
        missing = not os.path.exists(Config.model_IMU_path)
        self.imu_surrogate = missing if imu_surrogate is None else bool(imu_surrogate)
        self.quiet = quiet
        self.from_raw = bool(from_raw)
        self.use_graph = bool(use_graph)
        self.fused = bool(fused)    # False: chain the three drop-in modules per batch exactly as Demo_test.py:111-123 does
        self._graph = None
        self.graphed = False        # whether the last eval_model() replayed a captured step
        if self.from_raw:
            # the whole sample set is built on the device from raw sensor frames (one kernel launch); the surrogate
            # (R, t) are IMU_Net's training targets: R_R0R and the head joint (Train_IMU.py:127,138-139)
            from ...Util.Universal_Util.Dataset_sample import PosePC
            t0 = time.time()
            ds = PosePC(train=False, vis=True, batch_length=self.frame_no, device=self.device)
            b = ds.batch()
            torch.cuda.synchronize(self.device)
            self.build_seconds = time.time() - t0
            self.data, self.target, self.skl, self.imu = b["data"], b["key"], b["skl"], b["imu"]
            self.R_sur, self.t_sur = b["R"], b["key"][:, :, 20].contiguous()
        else:
            if not os.path.exists(Config.sample_frozen_path):
                raise FileNotFoundError(Config.sample_frozen_path)
            z = np.load(Config.sample_frozen_path)
            self.data = torch.from_numpy(z["data"]).float()
            self.target = torch.from_numpy(z["target"]).float()
            self.skl = torch.from_numpy(z["skl"]).float()
            self.imu = torch.from_numpy(z["imu"]).float()
            self.R_sur = torch.from_numpy(z["R_sur"]).float()
            self.t_sur = torch.from_numpy(z["t_sur"]).float()
            # page-locked host tensors: the per-batch host->device copies below are then asynchronous, so the host runs
            # ahead of the GPU instead of waiting for every staged copy (what limits the reference's one-snippet-per-call
            # setting is host time per call, not kernels)
            try:
                for name in ("data", "target", "skl", "imu", "R_sur", "t_sur"):
                    setattr(self, name, getattr(self, name).pin_memory())
            except RuntimeError:
                pass                     # pageable memory works too, only slower
        if missing and not quiet:
            print("IMU_Net checkpoint not found at %s: %s" % (
                Config.model_IMU_path,
                "using the ground-truth head pose as (R, t) [surrogate pin 2.6607 cm]" if self.imu_surrogate
                else "using seeded random IMU_Net weights (accuracy is meaningless)"))

    def eval_model(self):
        pipe, dev, bs = self.pipe, self.device, self.batch_size
        n = self.data.shape[0]
        sums = torch.zeros(SUMS_LEN, dtype=torch.float64, device=dev)
        h = pipe.handle
        on_dev = self.data.device == dev              # from_raw: the set already lives on the GPU
        state = {}                                    # zero initial LSTM states per batch size (read-only for the networks)
        # With IMU_Net in the chain the whole step is ONE library call (mmego_pipeline_forward: same kernels as the module
        # calls below, plus the fused head/decode/assembly/metrics tails); at the reference's own batch size of one snippet
        # that call is captured once as a CUDA graph and replayed per batch (GraphedStep).
        fused = self.fused and not self.imu_surrogate
        graph = None
        tc = time.time()
        if fused and self.use_graph and bs * self.frame_no <= self.GRAPH_MAX_SEQ and n >= bs:
            try:
                if self._graph is None or not self._graph.valid():
                    self._graph = None
                    self._graph = GraphedStep(pipe, bs, self.data.shape[1], self.data.shape[2], self.imu.shape[2])
                graph = self._graph
                graph.sums.zero_()
            except Exception as e:                    # capture refused (driver / allocator state): run the step eagerly
                if not self.quiet:
                    print("CUDA graph capture failed (%s); running eagerly" % (e,))
                torch.cuda.synchronize(dev)
                graph = self._graph = None
        self.graphed = graph is not None
        self.capture_seconds = time.time() - tc       # one-time set-up (like the weight upload), not part of `seconds`
        t0 = time.time()
        with torch.no_grad():
            for s in range(0, n, bs):
                e = min(n, s + bs)
                if graph is not None and e - s == bs:
                    graph.load(self.imu[s:e], self.data[s:e], self.skl[s:e], self.target[s:e])
                    graph.replay()
                    continue
                if fused:
                    pipe.forward(self.imu[s:e].to(dev, non_blocking=True),
                                 self.data[s:e].clone() if on_dev else self.data[s:e].to(dev, non_blocking=True),
                                 self.skl[s:e].to(dev, non_blocking=True), self.target[s:e].to(dev, non_blocking=True),
                                 sums, want_pred=False)
                    continue
                # forward transforms xyz in place: a batch copied from the host is already private, a device-resident set is cloned
                data = self.data[s:e].clone() if on_dev else self.data[s:e].to(dev, non_blocking=True)
                target = self.target[s:e].to(dev, non_blocking=True)
                skl = self.skl[s:e].to(dev, non_blocking=True)
                if self.imu_surrogate:
                    R = self.R_sur[s:e].to(dev, non_blocking=True)
                    t = self.t_sur[s:e].to(dev, non_blocking=True)
                else:
                    R, t = pipe.imu_net(self.imu[s:e].to(dev, non_blocking=True))
                b = e - s
                if b not in state:
                    state[b] = (torch.zeros(6, b, 64, device=dev), torch.zeros(6, b, 64, device=dev))
                h0, c0 = state[b]
                upper = pipe.upper_net(data, h0, c0, skl, R, t)[0]
                upper_l = upper.clone().detach()
                lower_l, _ = pipe.lower_net(upper_l, data, h0, c0, h0, c0, skl, R, t)
                h.assemble_metrics(upper_l, lower_l, target, sums, want_pred=False)
            if graph is not None:
                sums += graph.sums
            torch.cuda.synchronize(dev)
        self.seconds = time.time() - t0
        rep = report_from_sums(sums.cpu().numpy())
        eval_accu = rep["mpjpe_cm"] / 100.0
        accu_ll = rep["per_joint_cm"] / 100.0
        if not self.quiet:
            print("%d it in %.2f s = %.1f it/s" % (n, self.seconds, n / self.seconds))
            print('Average Joint Localization Error(cm): {}'.format(eval_accu * 100))
            print('Average UpperBody Joint Localization Error(cm): {}'.format(rep["upper_cm"]))
            print('Average LowerBody Joint Localization Error(cm): {}'.format(rep["lower_cm"]))
            print('Average Joint Rotation Error(°): {}'.format(rep["angle_deg"]))
            print('Per Joint Localization Error(cm): {}'.format(accu_ll * 100))
        self.report = rep
        return (rep["eval_loss"], rep["eval_loss_l"], eval_accu, rep["lower_cm"] / 100.0, accu_ll,
                rep["angle_bone_deg"])
