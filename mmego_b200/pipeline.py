"""The chained inference pass of Processor/Test/Demo_test.py:106-123 as one object: IMUNet -> UpperNet -> LowerNet ->
21-joint assembly (+ error sums), for device-resident batches (`forward`) and for host batches (`infer_host`).
Multi-GPU: one process per GPU, snippets sharded contiguously on dim 0, predictions all-gathered and error sums
all-reduced (`ShardedRunner`)."""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _capi, synth
from .Config.config import Config
from .Net.IMU_Net import IMUNet
from .Net.Lower_Net import LowerNet
from .Net.Upper_Net import UpperNet
from .engine import MMEgoError

SUMS_LEN = _capi.SUMS_LEN


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    return torch.load(path, map_location="cpu", weights_only=True)


def report_from_sums(sums) -> Dict[str, object]:
    """The five printed quantities of Demo_test.py:176-180 from the device-accumulated sums (layout: mmego_b200.h)."""
    s = np.asarray(sums, dtype=np.float64)
    F_ = s[43]
    return dict(frames=int(F_),
                mpjpe_cm=float(s[0:21].sum() / (F_ * 21) * 100.0),
                upper_cm=float(s[21] / (F_ * 15) * 100.0),
                lower_cm=float(s[22] / (F_ * 8) * 100.0),
                angle_deg=float((s[23:43] / F_).mean()),
                per_joint_cm=s[0:21] / F_ * 100.0,
                angle_bone_deg=s[23:43] / F_,
                eval_loss=float(s[44] / F_),
                eval_loss_l=np.asarray([s[44] / F_ / 8.0, s[45] / F_]))


class MMEgoPipeline:
    def __init__(self, device="cuda", imu_state: Optional[Mapping[str, torch.Tensor]] = None,
                 upper_state: Optional[Mapping[str, torch.Tensor]] = None,
                 lower_state: Optional[Mapping[str, torch.Tensor]] = None, body_index_mode: str = "ref"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MMEgoError("MMEgoPipeline needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.imu_net = IMUNet(15, 9, 512, 2, True, 0.1)
        self.upper_net = UpperNet()
        self.lower_net = LowerNet(64)
        self.imu_weights = "checkpoint"
        if imu_state is None:
            if os.path.exists(Config.model_IMU_path):
                imu_state = load_checkpoint(Config.model_IMU_path)
            else:
                imu_state = synth.imu_state_dict(0)        # the blob is missing from the reference mount
                self.imu_weights = "seeded-random(0)"
        self.imu_net.load_state_dict(imu_state)
        self.upper_net.load_state_dict(upper_state if upper_state is not None else load_checkpoint(Config.model_upper_path))
        self.lower_net.load_state_dict(lower_state if lower_state is not None else load_checkpoint(Config.model_lower_path))
        for m in (self.imu_net, self.upper_net, self.lower_net):
            m.to(self.device).eval()
        if body_index_mode not in ("ref", "per_snippet"):
            raise MMEgoError(f"body_index_mode must be 'ref' or 'per_snippet' (got {body_index_mode!r})")
        self.body_mode = _capi.BODY_REF if body_index_mode == "ref" else _capi.BODY_PER_SNIPPET
        # the drop-in modules follow the pipeline's mode ("ref" = the reference's initial_body[r % B] of ONE call)
        self.upper_net.body_index_mode = body_index_mode
        self.lower_net.body_index_mode = body_index_mode
        self.handle = None
        self._sync()

    def _sync(self):
        for m in (self.imu_net, self.upper_net, self.lower_net):
            self.handle = m._sync(self.device)

    def launch_count(self) -> int:
        return self.handle.launch_count()

    def forward(self, imu, data, skl, target=None, sums=None, b_offset: int = 0, B_global: Optional[int] = None,
                outs: Optional[Dict[str, torch.Tensor]] = None, want_pred: bool = True):
        """Device tensors in, pred [B,L,21,3] out.  `data` is transformed in place twice, as the reference does.
        When `target` and `sums` (float64[SUMS_LEN], device) are given the batch's error sums are ADDED to `sums`."""
        self._sync()
        return self.handle.pipeline_forward(imu, data, skl, target, sums, self.body_mode, b_offset, B_global, outs,
                                            want_pred)

    def infer_host(self, imu, data, skl, target=None, b_offset: int = 0, B_global: Optional[int] = None,
                   out_pred=None, out_sums=None):
        """Host tensors in, (pred, sums) host tensors out; copies are part of the call.  out_pred / out_sums: optional
        caller-owned pinned result buffers (reused across calls)."""
        self._sync()
        return self.handle.infer_host(imu, data, skl, target, self.body_mode, b_offset, B_global, out_pred, out_sums)


class GraphedStep:
    """One pipeline step at a FIXED shape captured as a CUDA graph: the latency form of `MMEgoPipeline.forward` for the
    reference's own operating point, one snippet per call (Processor/Test/Demo_test.py:61).  At that size a step is ~35
    launches of a few microseconds each and the host side (argument checks, tensor-map encoding, the launches themselves)
    is as long as the kernels; a replay is ONE driver call.  The caller copies a batch into the static input tensors
    (`imu`, `data`, `skl`, `target`) and calls `replay()`; `sums` accumulates over replays exactly as in `forward`, and
    `pred` is the static output (None when not wanted).  The graph holds the device addresses it was captured with: the
    workspace is kept alive here, and `valid()` tells whether the handle still holds the weights it was captured with."""

    def __init__(self, pipe: "MMEgoPipeline", B: int, L: int, N: int, n_imu: int, want_pred: bool = False,
                 with_target: bool = True):
        dev = pipe.device
        pipe._sync()
        self.pipe, self.handle = pipe, pipe.handle
        self.imu = torch.zeros(B, L, n_imu, 15, device=dev)
        self.data = torch.zeros(B, L, N, 6, device=dev)
        self.skl = torch.zeros(B, 20, 3, device=dev)
        self.target = torch.zeros(B, L, 21, 3, device=dev) if with_target else None
        self.sums = torch.zeros(SUMS_LEN, dtype=torch.float64, device=dev) if with_target else None
        self.pred = None
        h = self.handle

        def step():
            return h.pipeline_forward(self.imu, self.data, self.skl, self.target, self.sums, pipe.body_mode, 0, None, None,
                                      want_pred)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):            # warm-up off the capture: lazy allocations, function attributes, workspace size
            step()
            step()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self._ws = h._ws                         # the captured launches point into this workspace: keep it alive
        self._owner = dict(h.weights_owner)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.pred = step()
        if self.sums is not None:
            self.sums.zero_()                    # the warm-up steps added to it
        self.launches_per_replay = None

    def valid(self) -> bool:
        self.pipe._sync()
        return all(self.handle.weights_owner.get(k) is v for k, v in self._owner.items()) and self.handle._ws is self._ws

    def load(self, imu, data, skl, target=None):
        """Copies one batch (host or device tensors of the captured shape) into the static inputs, asynchronously."""
        self.imu.copy_(imu, non_blocking=True)
        self.data.copy_(data, non_blocking=True)
        self.skl.copy_(skl, non_blocking=True)
        if target is not None and self.target is not None:
            self.target.copy_(target, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.pred


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the snippet dimension; the first B % world ranks take one extra snippet."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedRunner:
    """Data-parallel wrapper: each rank runs `step_fn` on its contiguous shard; predictions are all-gathered and the
    float64 error sums all-reduced.  Snippets are independent, so this is the only communication of the path.

    `run` is the synchronous form (collectives at the end of the step on the compute stream).  `submit` / `collect` split
    it: `submit` runs the shard's step and ENQUEUES the collectives on a side stream (after an event on the compute
    stream), so the gather of step i travels over NVLink while step i+1 computes; `collect` makes the current stream wait
    for the oldest outstanding collective and returns its (pred, sums).  Two result slots are kept, so at most two steps
    may be outstanding.  Measured on 8 B200s: the overlapped form is SLOWER for this pipeline (weak efficiency 0.936 vs
    0.968): NCCL's CTAs spin on a few SMs while they wait for the slowest rank, and the persistent LSTM kernel partitions
    its work statically over all 148 SMs, so the slowed SMs set the time of every launch -- `bench.py` defaults to `run`.  With equal shards the gathered buffer already IS the final [B, L, 21, 3] layout (rank-major =
    snippet order): no concatenation copy.  `collective_ms()` returns the device time spent in the collectives."""

    def __init__(self, step_fn, world: int, rank: int, group=None):
        self.step_fn, self.world, self.rank, self.group = step_fn, world, rank, group
        self._slots = [None, None]
        self._pending = []
        self._side = None
        self._events = []

    # ------------------------------------------------------------------------------------------ helpers
    def _sizes(self, B):
        return [shard_bounds(B, self.world, r) for r in range(self.world)]

    def _gather(self, B, pred, sums, slot):
        """Enqueues all-gather(pred) and all-reduce(sums) on the CURRENT stream; returns (gathered view, sums)."""
        import torch.distributed as dist
        sizes = self._sizes(B)
        rows = max(h - l for l, h in sizes)                    # uneven shards: pad to the largest, trim after
        mine = pred.contiguous()
        if mine.shape[0] < rows:
            mine = torch.cat((mine, mine.new_zeros((rows - mine.shape[0],) + tuple(mine.shape[1:]))), dim=0)
        buf = self._slots[slot]
        shape = (self.world * rows,) + tuple(pred.shape[1:])
        if buf is None or tuple(buf.shape) != shape or buf.dtype != pred.dtype or buf.device != pred.device:
            buf = torch.empty(shape, dtype=pred.dtype, device=pred.device)
            self._slots[slot] = buf
        dist.all_gather_into_tensor(buf, mine, group=self.group)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        if all(h - l == rows for l, h in sizes):
            return buf, sums                                   # already the final layout
        return torch.cat([buf[r * rows:r * rows + (h - l)] for r, (l, h) in enumerate(sizes)], dim=0), sums

    # ------------------------------------------------------------------------------------------ synchronous
    def run(self, B: int, *step_args):
        lo, hi = shard_bounds(B, self.world, self.rank)
        pred, sums = self.step_fn(lo, hi, B, *step_args)      # pred [hi-lo, L, 21, 3], sums float64[SUMS_LEN]
        if self.world == 1:
            return pred, sums
        if pred.is_cuda:                                       # device time of the collectives (incl. waiting for the slowest rank)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = self._gather(B, pred, sums, 0)
            e1.record()
            self._events.append((e0, e1))
            return out
        return self._gather(B, pred, sums, 0)

    # ------------------------------------------------------------------------------------------ pipelined
    def submit(self, B: int, *step_args):
        lo, hi = shard_bounds(B, self.world, self.rank)
        pred, sums = self.step_fn(lo, hi, B, *step_args)
        if self.world == 1:
            self._pending.append((pred, sums, None))
            return
        if len(self._pending) >= 2:
            raise MMEgoError("ShardedRunner: collect() a step before submitting a third one (two result slots)")
        if pred.is_cuda:
            cur = torch.cuda.current_stream(pred.device)
            if self._side is None:
                self._side = torch.cuda.Stream(pred.device)
            sums = sums.clone()                                # (compute stream) the caller may zero its accumulator next step
            ready = torch.cuda.Event()
            ready.record(cur)
            slot = self._seq = (getattr(self, "_seq", -1) + 1) % 2
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self._side)
                pred.record_stream(self._side)
                sums.record_stream(self._side)
                out = self._gather(B, pred, sums, slot)
                e1.record(self._side)
            self._events.append((e0, e1))
            self._pending.append((out[0], out[1], e1))
        else:                                                  # CPU tensors (gloo tests): no streams, same bookkeeping
            slot = self._seq = (getattr(self, "_seq", -1) + 1) % 2
            out = self._gather(B, pred, sums.clone(), slot)
            self._pending.append((out[0], out[1], None))

    def collect(self):
        pred, sums, done = self._pending.pop(0)
        if done is not None:
            torch.cuda.current_stream(pred.device).wait_event(done)
        return pred, sums

    def collective_ms(self) -> float:
        total = 0.0
        for e0, e1 in self._events:
            e1.synchronize()
            total += e0.elapsed_time(e1)
        self._events = []
        return total
