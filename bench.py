#!/usr/bin/env python
"""bench.py -- mmEgo inference frames/s on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL): snippets are independent, so every rank runs the whole pipeline
on its own `--batch` snippets (weak scaling) and the only communication is the all-gather of predictions and the
all-reduce of the error sums, inside the timed region.

A step = one pass of the full pipeline (IMU_Net -> Upper_Net -> Lower_Net -> 21-joint assembly + error sums) over the
batch, through the C ABI (`mmego_pipeline_forward`).  `value` times it with inputs resident in HBM; `e2e` times
`mmego_infer_host` with pinned HOST buffers (H2D of the inputs and D2H of predictions + sums inside the timed region).
Inputs per step are 350 MB at B=4096 -- larger than the 126 MB L2 -- so consecutive steps do not hit in L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mmEgo inference frames/s (IMU_Net+Upper_Net+Lower_Net/GCN+decode, fp32 parity mode)"
KERNEL_NAMES = {0: "gemm_ffma_kernel<128,EPI_LSTM> (fp32 FFMA GEMM + fused LSTM cell)",
                1: "lstm_tc_step_kernel<3> (tcgen05 fp16x3 split-precision GEMM, TMEM partial sums drained to fp32 registers, fused LSTM cell)",
                2: "lstm_tc_step_kernel<1> (tcgen05 fp16 GEMM, fused LSTM cell)"}
UNIT = "frames/s"
_REAL_STDOUT = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)
L, N_PTS, N_IMU = 20, 128, 20
H = 512
# algorithmic FLOPs (2*MAC) per frame, SURVEY.md section 8(d)
FLOPS_PER_FRAME = dict(imu=444_962_816, upper=2_150_922, lower=9_267_786)
FLOPS_PER_FRAME["total"] = sum(FLOPS_PER_FRAME.values())


def lstm_step_flops(seqs: int, in_features: int) -> float:
    """One timestep launch (both directions) of an H=512 LSTM layer: 2 dirs * 2 * M * 4H * (In + H)."""
    return 2 * 2.0 * seqs * 4 * H * (in_features + H)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic_from_profiles(kernel_substr="lstm_tc_step_kernel<3"):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, read from the NEWEST
    committed `ncu --set full` summary under profiles/ (scripts/ncu_summary.py output) that holds a capture of it; among
    that file's launches the one with the most traffic is the rnn_fast layer-1 step (M = 40,960, K = 1536)."""
    import glob
    import re
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.txt")):
        m = re.match(r"r(\d+)", os.path.basename(path))
        rnd = int(m.group(1)) if m else 0
        top, cur = None, None
        for ln in open(path, errors="replace"):
            if ln.startswith("Kernel Name"):
                cur = {"hit": kernel_substr in ln, "rd": None, "wr": None}
            elif cur is not None and cur["hit"]:
                f = ln.split()
                if ln.startswith("dram__bytes_read.sum ") or ln.startswith("dram__bytes_write.sum "):
                    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[-1], None)
                    if mult is not None:
                        cur["rd" if "read" in f[0] else "wr"] = float(f[1].replace(",", "")) * mult
                    if cur["rd"] is not None and cur["wr"] is not None:
                        tot = cur["rd"] + cur["wr"]
                        if top is None or tot > top:
                            top = tot
                        cur = None
        if top is not None:
            key = (rnd, os.path.getmtime(path), os.path.basename(path))
            if best is None or key > best[0]:
                best = (key, top, os.path.relpath(path, ROOT))
    return (best[1], best[2]) if best else (None, None)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def load_checkpoints():
    import torch
    base = os.path.join(ROOT, "Resource", "Pretrained_model")
    up = torch.load(os.path.join(base, "Upper_Net", "epoch451_batch20frame20lr3e-05.pth"), map_location="cpu", weights_only=True)
    lo = torch.load(os.path.join(base, "Lower_Net", "epoch161_batch20frame20lr0.0003.pth"), map_location="cpu", weights_only=True)
    return up, lo


class ReferenceChain:
    """The reference's OWN nn.Module classes (vendored, unmodified, into the git-ignored baseline/_ref/ by
    scripts/vendor_reference.py), chained as Processor/Test/Demo_test.py:106-123 on the host cores."""
    kind = "reference"

    def __init__(self):
        import types
        import torch
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.animation", "seaborn", "imageio",
                     "imageio.v2", "mpl_toolkits", "mpl_toolkits.mplot3d"):
            sys.modules.setdefault(name, types.ModuleType(name))       # plotting imports of Utils.py; absent in this image
        sys.path.insert(0, REF_DIR)
        from Net.IMU_Net import IMUNet          # noqa: E402  (baseline/_ref)
        from Net.Lower_Net import LowerNet      # noqa: E402
        from Net.Upper_Net import UpperNet      # noqa: E402
        from Config.config import Config        # noqa: E402
        from mmego_b200 import synth
        self.torch, self.Config = torch, Config
        up_sd, lo_sd = load_checkpoints()
        self.imu = IMUNet(15, 9, 512, 2, True, 0.1)                     # Demo_test.py:54
        self.imu.load_state_dict(synth.imu_state_dict(0))               # the checkpoint blob is missing upstream
        self.upper, self.lower = UpperNet(), LowerNet(64)
        self.upper.load_state_dict(up_sd)
        self.lower.load_state_dict(lo_sd)
        for m in (self.imu, self.upper, self.lower):
            m.eval()

    def __call__(self, imu, data, skl):
        torch, Config = self.torch, self.Config
        B, L = data.shape[:2]
        data = data.clone()                                             # the nets transform the cloud in place
        h0 = torch.zeros((6, B, 64), dtype=torch.float32)
        c0 = torch.zeros((6, B, 64), dtype=torch.float32)
        R_p, t_p = self.imu(imu)                                        # Demo_test.py:111
        R, t = R_p.clone().detach(), t_p.clone().detach()
        upper, _, _, _, _ = self.upper(data, h0, c0, skl, R, t)         # :114
        upper_l = upper.clone().detach()
        lower_l, _ = self.lower(upper_l, data, h0, c0, h0, c0, skl, R, t)   # :118
        pred = torch.zeros((B, L, 21, 3), dtype=torch.float32)
        pred[:, :, Config.upper_joint_map, :] = upper_l                 # :121-123
        pred[:, :, Config.lower_joint_map, :] = lower_l
        return pred


class PortChain:
    """Fallback when baseline/_ref/ is absent: the oracle port of the same chain (oracle/mmego_oracle.py)."""
    kind = "port"

    def __init__(self):
        from mmego_b200 import synth
        from oracle import mmego_oracle as O
        self.O, self.imu_sd = O, synth.imu_state_dict(0)
        self.up, self.lo = load_checkpoints()

    def __call__(self, imu, data, skl):
        return self.O.pipeline(self.imu_sd, self.up, self.lo, imu, data, skl)["pred"]


def make_cpu_chain():
    if os.path.isdir(os.path.join(REF_DIR, "Net")):
        return ReferenceChain()
    return PortChain()


def cpu_reference_pass(snippets: int, steps: int, warmup: int, batch1_snippets: int = 0):
    """Times the reference's CPU path on a bounded sample of the synthetic workload: `snippets` per step in ONE batched call
    (B = 128 is the best CPU throughput observed), and optionally `batch1_snippets` snippets one at a time
    (B = 1, the reference's own DataLoader setting, Processor/Test/Demo_test.py:61).  All host threads, fp32, no_grad."""
    import torch
    from mmego_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    chain = make_cpu_chain()
    sb = synth.batch(snippets, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            chain(sb["imu"], sb["data"], sb["skl"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        b1 = None
        if batch1_snippets > 0:
            chain(sb["imu"][:1], sb["data"][:1], sb["skl"][:1])
            t0 = time.perf_counter()
            for j in range(batch1_snippets):
                chain(sb["imu"][j:j + 1], sb["data"][j:j + 1], sb["skl"][j:j + 1])
            dt1 = (time.perf_counter() - t0) / batch1_snippets
            b1 = {"value": L / dt1, "unit": UNIT, "ms_per_snippet": dt1 * 1e3, "snippets": batch1_snippets,
                  "what": "batch = 1 snippet per call (the reference's own setting, Demo_test.py:61)"}
    dt = sum(times) / len(times)
    return dict(fps=snippets * L / dt, dt=dt, best_fps=snippets * L / min(times), cores=cores, kind=chain.kind, batch1=b1,
                cpu=cpu_model_name())


def cpu_sample_text(r, snippets, steps, warmup):
    what = ("the reference's own IMUNet/UpperNet/LowerNet classes (baseline/_ref, unmodified) chained as Demo_test.py:111-123"
            if r["kind"] == "reference" else "oracle port of the reference's PyTorch-CPU path (baseline/_ref absent)")
    return (f"{snippets} snippets ({snippets * L} frames) of the same synthetic workload per step in one batched call, "
            f"{warmup} warm-up + {steps} timed steps; {what}; PyTorch CPU fp32, torch threads = {r['cores']} ({r['cpu']})")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""          # the reference's Config picks cuda:0 when it sees one; this arm is the CPU path
    snippets = args.cpu_snippets
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference_pass(snippets, steps, warmup, batch1_snippets=16)
    fps = r["fps"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["dt"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"full pipeline, B={args.batch}, L={L}, N={N_PTS}, n_imu={N_IMU} (config 3 of BASELINE.json)",
                   "note": f"each step is a bounded sample of that workload: {snippets} snippets in one batched CPU call"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": cpu_sample_text(r, snippets, steps, warmup), "best_step_value": r["best_fps"],
                         "batch1": r["batch1"]},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
SURROGATE_PIN_CM = 2.660650576      # reference classes, 835 sample snippets, IMU_Net's training targets as (R, t)


def config1_record(dev):
    """Config 1 of BASELINE.json (the 835 real sample snippets through the drop-in `MMEgo().eval_model()`):
    MPJPE with the surrogate head pose (the IMU_Net checkpoint is missing upstream; pin = the reference's own classes),
    and the wall-clock it/s of the full chain INCLUDING IMU_Net at batch 167 and at the reference's batch 1."""
    import numpy as np
    import torch
    from mmego_b200.Config.config import Config
    from mmego_b200.Processor.Test.Demo_test import MMEgo
    if not os.path.exists(Config.sample_frozen_path):
        return None
    m = MMEgo(batch_size=167, device=dev, imu_surrogate=True, quiet=True)
    m.eval_model()
    rec = {"workload": "835 sample snippets x 20 frames (Resource/Sample_data, frozen tensors), MMEgo().eval_model()",
           "mpjpe_surrogate_cm": m.report["mpjpe_cm"], "mpjpe_surrogate_pin_cm": SURROGATE_PIN_CM,
           "upper_cm": m.report["upper_cm"], "lower_cm": m.report["lower_cm"], "angle_deg": m.report["angle_deg"],
           "pin_note": "IMU_Net's training targets as (R, t); README's 3.893 cm needs the missing IMU_Net checkpoint"}
    for bs, key in ((167, "batch167"), (1, "batch1")):
        mm = MMEgo(batch_size=bs, device=dev, imu_surrogate=False, quiet=True)
        mm.eval_model()                      # warm-up (weight packing, allocator)
        mm.eval_model()
        n = int(mm.data.shape[0])
        rec[key] = {"it_per_s": n / mm.seconds, "frames_per_s": n * 20 / mm.seconds, "ms_per_snippet": mm.seconds / n * 1e3,
                    "cuda_graph_replay": bool(mm.graphed),
                    "what": f"full chain incl. IMU_Net (seeded stand-in weights), pinned host tensors in, batch = {bs} snippet(s) "
                            "per call" + (", the step captured once as a CUDA graph and replayed per call" if mm.graphed else "")}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="snippets per GPU (config 3: 4096)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-snippets", type=int, default=128, help="snippets per CPU-baseline step (bounded sample; 128 = best CPU batch)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-config1", action="store_true", help="skip the 835-sample-snippet record (config 1)")
    ap.add_argument("--collective", default="sync", choices=["sync", "async"],
                    help="N>1: sync = all-gather / all-reduce at the end of every step on the compute stream (default); async = "
                         "on a side stream under the next step's compute (measured slower at N=8: the spinning NCCL CTAs share "
                         "SMs with the persistent, statically partitioned LSTM kernel, profiles/r02_scale_*.json)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = --batch snippets PER GPU (default, config 3 per GPU); strong = --batch snippets in total, "
                         "split contiguously over the GPUs (config 4 of BASELINE.json)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (repeatable)")
    ap.add_argument("--no-half", action="store_true", help="skip the extra pass in single-pass fp16 mode")
    ap.add_argument("--imu-gemm", type=int, default=None, help="0 fp32 FFMA, 1 tcgen05 fp16x3, 2 tcgen05 fp16 (default: library default)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything else written to fd 1 while the bench runs (NCCL's version banner is
    # printed from C) is sent to stderr, and the line is written to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    # the JSON line must be the only thing on stdout: NCCL's version banner (NCCL_DEBUG=VERSION) goes there too
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    from mmego_b200 import _capi, synth
    from mmego_b200.pipeline import SUMS_LEN, MMEgoPipeline, ShardedRunner, report_from_sums

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong" and world > 1
    if strong:
        from mmego_b200.pipeline import shard_bounds
        Bg = args.batch                              # strong scaling: the SAME global batch, split contiguously on dim 0
        lo_, hi_ = shard_bounds(Bg, world, rank)
        B = hi_ - lo_
    else:
        B = args.batch
        Bg = B * world                               # weak scaling: every rank owns B snippets
        lo_ = 0
    pipe = MMEgoPipeline(dev, imu_state=None)
    if args.imu_gemm is not None:
        pipe.handle.set_option("imu_gemm", args.imu_gemm)
    for kv in args.opt:
        k, v = kv.split("=")
        pipe.handle.set_option(k, int(v))
    # this rank's shard of the global synthetic batch (seed depends on the rank; same distribution)
    if strong:      # every rank generates the global batch and keeps its slice; skeletons stay global (initial_body[r % B_global])
        sb = synth.batch(Bg, L=L, N=N_PTS, n_imu=N_IMU, seed=1234)
        sb = dict(imu=sb["imu"][lo_:hi_].contiguous(), data=sb["data"][lo_:hi_].contiguous(), skl=sb["skl"])
    else:
        sb = synth.batch(B, L=L, N=N_PTS, n_imu=N_IMU, seed=1234 + rank)
    fwd_kw = dict(b_offset=lo_, B_global=Bg) if strong else {}
    imu_h, data_h, skl_h = sb["imu"].pin_memory(), sb["data"].pin_memory(), sb["skl"].pin_memory()
    imu_d, data0_d, skl_d = imu_h.to(dev), data_h.to(dev), skl_h.to(dev)
    data_d = torch.empty_like(data0_d)
    sums_d = torch.zeros(SUMS_LEN, dtype=torch.float64, device=dev)
    # synthetic ground truth: first prediction + 3 cm noise
    data_d.copy_(data0_d)
    pred0 = pipe.forward(imu_d, data_d, skl_d, **fwd_kw)
    target_h = synth.target_like(pred0, seed=99 + rank).contiguous().pin_memory()
    target_d = target_h.to(dev)
    torch.cuda.synchronize()

    def step_fn(lo, hi, Bglobal):
        # the in-place Transform2H mutates the cloud: every step starts from a fresh copy of the resident input
        data_d.copy_(data0_d)
        sums_d.zero_()
        pred = pipe.forward(imu_d, data_d, skl_d, target_d, sums_d, **fwd_kw)
        return pred, sums_d

    runner = ShardedRunner(lambda lo, hi, Bglobal: step_fn(lo, hi, Bglobal), world, rank)

    def one_step():
        # every rank holds exactly its own shard of the Bg snippets (weak: B each; strong: shard_bounds of --batch)
        return runner.run(Bg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_pass(steps, warmup):
        """W untimed + K timed steps; returns (ms max over ranks, launches, per-span profile, clocks, sums)."""
        for _ in range(warmup):
            pred, sums = one_step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        pipe.handle.profile_begin()
        n0 = pipe.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            runner.collective_ms()      # drop the warm-up steps' events: the figure below covers the timed steps only
        barrier()
        e0.record()
        if world > 1 and args.collective == "async":
            # the gather / all-reduce of step i runs on a side stream while step i+1 computes (two result slots):
            # no rank waits inside a step for the slowest GPU of the box; every step's result is collected
            for i in range(steps):
                runner.submit(Bg)
                if i >= 1:
                    pred, sums = runner.collect()
            pred, sums = runner.collect()
        else:
            for _ in range(steps):
                pred, sums = one_step()
        e1.record()
        barrier()
        ms_ = e0.elapsed_time(e1)
        launches_ = pipe.launch_count() - n0
        coll_ms_ = runner.collective_ms() / steps if world > 1 else 0.0
        prof_ = pipe.handle.profile_read()
        pipe.handle.profile_end()
        clocks_ = sampler.stop() if rank == 0 else None
        t_ms = torch.tensor([ms_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        timed_pass.collective_ms = coll_ms_
        return float(t_ms.item()), launches_, prof_, clocks_, sums

    mode = args.imu_gemm if args.imu_gemm is not None else 1
    ms, launches, prof, clocks, sums = timed_pass(args.steps, args.warmup)
    collective_ms = timed_pass.collective_ms
    ms_per_step = ms / args.steps
    # N > 1: every rank's own pace on the same work WITHOUT the per-step collectives (a few untimed-by-the-metric local
    # steps): shows how much of the loss against N x the 1-GPU figure is the spread between the box's power-capped GPUs
    # (the lock-stepped run moves at the pace of the slowest one)
    rank_free_ms = None
    if world > 1:
        k = max(2, min(args.steps, 5))
        step_fn(0, 0, Bg)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k):
            step_fn(0, 0, Bg)      # the arguments are unused: the rank's shard is already resident
        f1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([f0.elapsed_time(f1) / k], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_free_ms = [round(float(x.item()), 3) for x in allr]
    frames = Bg * L
    value = frames / (ms_per_step * 1e-3)
    rep = report_from_sums(sums.cpu().numpy())

    # ---- end to end through the host-buffer entry point (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        # caller-owned pinned result buffers, as a serving loop would keep them (a fresh 20 MB pinned allocation per call
        # costs ~10 ms)
        out_pred = torch.empty(B, L, 21, 3, dtype=torch.float32).pin_memory()
        out_sums = torch.zeros(SUMS_LEN, dtype=torch.float64).pin_memory()
        for _ in range(2):
            pipe.infer_host(imu_h, data_h, skl_h, target_h, out_pred=out_pred, out_sums=out_sums, **fwd_kw)
        barrier()
        k2 = max(3, min(args.steps, 8))
        step_s = []
        t0 = time.perf_counter()
        for _ in range(k2):
            ts = time.perf_counter()
            pred_h, sums_h = pipe.infer_host(imu_h, data_h, skl_h, target_h, out_pred=out_pred, out_sums=out_sums, **fwd_kw)   # synchronous on return
            step_s.append(time.perf_counter() - ts)
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = sum(t.numel() * t.element_size() for t in (imu_h, data_h, skl_h, target_h))
        d2h = pred_h.numel() * 4 + SUMS_LEN * 8
        # host-to-device bandwidth of this box (one pinned 256 MB copy), to read the e2e figure against
        probe = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
        probe_d = torch.empty_like(probe, device=dev)
        probe_d.copy_(probe, non_blocking=True)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        probe_d.copy_(probe, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = probe.numel() * 4 / (time.perf_counter() - tp) / 1e9
        e2e = {"value": frames / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": float(dt.item()) * 1e3, "steps": k2,
               "step_ms_rank0": [round(x * 1e3, 2) for x in step_s], "h2d_gbs_rank0": round(h2d_gbs, 1),
               "api": "mmego_infer_host (pinned host buffers)"}

    half = None
    if mode == 1 and not args.no_half:
        pipe.handle.set_option("imu_gemm", 2)
        half = timed_pass(args.steps, 2)
        pipe.handle.set_option("imu_gemm", 1)

    if rank == 0:
        peaks = measured_peaks()
        # dominant kernel: the H=512 LSTM timestep launch of rnn_fast (2 layers x n_imu steps per IMU chunk, both
        # directions per launch, M = chunk*L sequences); rnn_slow's launches (M = chunk) are timed separately
        chunk = min(B, 2048)
        nchunks = (B + chunk - 1) // chunk
        fl_fast = 0.0
        for c in range(nchunks):
            bc = min(chunk, B - c * chunk)
            fl_fast += N_IMU * (lstm_step_flops(bc * L, H) + lstm_step_flops(bc * L, 2 * H))
        peak = peaks["bf16_sustained"]
        # DRAM bytes of one rnn_fast layer-1 step launch at M = 40,960: read from the newest ncu --set full summary in profiles/
        traffic_bytes, traffic_src = ncu_traffic_from_profiles()
        NCU_TRAFFIC = {1: traffic_bytes, 2: None, 0: None}

        def lstm_roofline(prof_, ms_, steps_, mode_):
            lst = prof_.get("imu.lstm_fast", dict(ms=0.0, launches=0))
            per_launch_flops = fl_fast * steps_ / max(1, lst["launches"])
            per_launch_ms = lst["ms"] / max(1, lst["launches"])
            ach = per_launch_flops / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
            passes = 3 if mode_ == 1 else 1
            slow = prof_.get("imu.lstm_slow", dict(ms=0.0))
            return {"bound": "tensor", "kernel": "H=512 LSTM timestep launch (rnn_fast): " + KERNEL_NAMES[mode_],
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": NCU_TRAFFIC.get(mode_) if B >= 2048 else None,
                    "traffic_note": f"ncu dram read+write of the rnn_fast launch with the most traffic in {traffic_src}: a second-layer timestep (M = 40,960 sequences, K = 1536); its algorithmic bytes are 1.06e9",
                    "peak_source": peaks["source"] + ", sustained dense bf16 (kernel timed inside a long step)",
                    "avg_launch_ms": per_launch_ms, "launches": lst["launches"], "flops_per_launch": per_launch_flops,
                    "mma_passes": passes, "tensor_pipe_tflops_issued": ach * passes,
                    "note": ("fp32-grade results need 3 fp16 tensor-core products per algorithmic multiply; "
                             "`achieved` counts algorithmic FLOPs once, `tensor_pipe_tflops_issued` is the MMA work done")
                            if passes == 3 else "single-pass fp16",
                    "share_of_step": (lst["ms"] + slow["ms"]) / ms_ if ms_ > 0 else None}

        roofline = lstm_roofline(prof, ms, args.steps, mode)
        stage_ms = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items()}

        def stage_rooflines(prof_, steps_, mode_):
            """Memory-bound stages: algorithmic HBM bytes per step / CUDA-event span of the stage, against the measured
            copy bandwidth; the mma.sync stages (point encoders, H=64 LSTMs) as algorithmic TFLOP/s for reference."""
            planes = 2 if mode_ == 1 else (1 if mode_ == 2 else 0)
            F_ = B * L
            rows = F_ * N_IMU
            act = 2 * planes if planes else 4                       # bytes per activation element (fp16 planes / fp32)
            hbm_bytes = {
                "imu.fc1": rows * (15 * 4 + 512 * act),
                "imu.pool": F_ * (N_IMU * 1024 * act + 1024 * act),
                "imu.decode": F_ * (1024 * act + 48),
                "upper.point": F_ * (N_PTS * 6 * 4 + N_PTS * 3 * 4 + 64 * 4 + 48),
                # fused tails (heads_mma.cu): head input in; joints (+ pred) out; R, t; lower also reads upper_l and target
                "upper.head_decode": F_ * (128 * 4 + 48 + 45 * 4),
                "lower.head_decode": F_ * ((128 + 45) * 4 + 48 + 24 * 4 + 45 * 4 + 63 * 4 + 63 * 4),
            }
            flops = {"upper.point": F_ * 1.57e6, "lower.frame": F_ * 1.40e6, "small_lstm": F_ * (0.524e6 + 0.655e6),
                     "lower.gcn": F_ * 7.18e6}
            out = {}
            for name, v in prof_.items():
                ms_s = v["ms"] / steps_
                if ms_s <= 0:
                    continue
                e = {"ms": round(ms_s, 3)}
                if name in hbm_bytes:
                    gbs = hbm_bytes[name] / (ms_s * 1e-3) / 1e9
                    e.update({"hbm_bytes": hbm_bytes[name], "gbs": round(gbs, 1)})
                    if name.endswith("head_decode"):
                        # the former HBM-bound decode / assembly / metrics kernels no longer exist: they run inside the head
                        # GEMM kernel (heads_mma.cu), whose time is the mma.sync chain + the per-frame decode, not its bytes
                        e["bound"] = "fused into the head GEMM (mma.sync chain + per-frame decode); bytes are the compulsory ones"
                    else:
                        e["hbm_frac"] = round(gbs / peaks["hbm_gbs"], 3)
                if name in flops:
                    e["algorithmic_tflops"] = round(flops[name] / (ms_s * 1e-3) / 1e12, 2)
                if len(e) > 1:
                    out[name] = e
            return out
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "f32 (fp16x3 split products, fp32 accumulate)", 2: "f16 (fp32 accumulate)"}[mode],
            "data": "synthetic",
            "config": {"workload": (f"full pipeline, B={Bg} snippets split over {world} GPUs, L={L}, N={N_PTS}, n_imu={N_IMU} "
                                    "(config 4 of BASELINE.json)" if strong else
                                    f"full pipeline, B={B} snippets per GPU, L={L}, N={N_PTS}, n_imu={N_IMU} "
                                    "(config 3 of BASELINE.json)") + "; IMU_Net weights: " + pipe.imu_weights,
                       "global_batch": Bg, "frames_per_step": frames, "parallelism": f"dp{world}",
                       "l2": "inputs per step (350 MB per GPU) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": launches,
            "collective_ms_per_step": round(collective_ms, 4) if world > 1 else None,
            "collective": args.collective if world > 1 else None,
            "rank_free_ms_per_step": rank_free_ms,
            "collective_note": ("all-gather of pred (into the final layout) + all-reduce of 46 float64 sums per step; device time "
                                "between CUDA events around them on rank 0, including the wait for the slowest rank") if world > 1 else None,
            "algorithmic_tflops": FLOPS_PER_FRAME["total"] * frames / (ms_per_step * 1e-3) / 1e12,
            "stage_ms_per_step": stage_ms,
            "mpjpe_vs_synthetic_target_cm": rep["mpjpe_cm"],
            "mpjpe_surrogate_cm": None, "config1": None,
            "roofline": roofline,
            "stages": stage_rooflines(prof, args.steps, mode),
        }
        if half:
            ms2, _, prof2, _, _ = half
            line["half_mode"] = {
                "what": "same workload with imu_gemm=2 (single-pass fp16 tensor-core LSTM); tolerance vs the fp32 oracle: "
                        "R 2e-3, t 1e-4 m, joints 3e-3 m with the seeded stand-in weights (tests/_parity.py IMU_MODE_TOL)",
                "value": frames / (ms2 / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                "roofline": lstm_roofline(prof2, ms2, args.steps, 2),
                "stage_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in prof2.items()}}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_config1:
            c1 = config1_record(dev)
            if c1:
                line["config1"] = c1
                line["mpjpe_surrogate_cm"] = c1["mpjpe_surrogate_cm"]
                line["mpjpe_surrogate_pin_cm"] = SURROGATE_PIN_CM
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_pass(args.cpu_snippets, 3, 1, batch1_snippets=8)
            line["cpu_baseline"] = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                    "sample": cpu_sample_text(r, args.cpu_snippets, 3, 1), "batch1": r["batch1"]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
