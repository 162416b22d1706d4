"""ctypes binding of libmmego_b200.so (C ABI declared in include/mmego_b200.h).

PyTorch is used for device memory and streams only: every call below passes raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream handle to the library.  There is no CPU
fallback: if the library has not been built, or a tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Mapping, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmmego_b200.so")

NET_IMU, NET_UPPER, NET_LOWER = 0, 1, 2
STAGE_IMU, STAGE_UPPER, STAGE_LOWER, STAGE_GCN, STAGE_PIPELINE = 0, 1, 2, 3, 4
BODY_REF, BODY_PER_SNIPPET = 0, 1
SUMS_LEN = 46

_vp, _i, _ll, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/mmego_b200.h one to one
SIGNATURES = {
    "mmego_abi_version": (_i, []),
    "mmego_create": (_i, [C.POINTER(_vp), _i]),
    "mmego_destroy": (_i, [_vp]),
    "mmego_last_error": (C.c_char_p, [_vp]),
    "mmego_set_option": (_i, [_vp, C.c_char_p, _ll]),
    "mmego_set_weights": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_ll), _i]),
    "mmego_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i, _i]),
    "mmego_imu_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "mmego_upper_forward": (_i, [_vp] + [_vp] * 11 + [_i] * 6 + [_vp, _sz, _vp]),
    "mmego_lower_forward": (_i, [_vp] + [_vp] * 7 + [_i] * 6 + [_vp, _sz, _vp]),
    "mmego_gcn_extract_feature": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "mmego_transform2h": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "mmego_transform2r": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "mmego_assemble_metrics": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "mmego_pipeline_forward": (_i, [_vp] + [_vp] * 10 + [_i] * 7 + [_vp, _sz, _vp]),
    "mmego_infer_host": (_i, [_vp] + [_vp] * 6 + [_i] * 7),
    "mmego_build_snippets": (_i, [_vp, _vp, _vp, _vp, C.c_uint, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mmego_debug_tap": (_i, [_vp, C.c_char_p, _vp, _sz]),
    "mmego_launch_count": (_ll, [_vp]),
    "mmego_profile_begin": (_i, [_vp]),
    "mmego_profile_read": (_i, [_vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(_ll), C.POINTER(_ll)]),
    "mmego_profile_end": (_i, [_vp]),
    "mmego_debug_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
}

PROFILE_SPANS = ("imu.fc1", "imu.lstm_fast", "imu.lstm_slow", "imu.pool", "imu.decode", "upper.point", "small_lstm", "upper.head_decode",
                 "lower.gcn", "lower.frame", "lower.head_decode", "assemble_metrics", "imu.resident",
                 "gcn.agg0", "gcn.gconv0", "gcn.tconv0", "gcn.agg1", "gcn.gconv1", "gcn.tconv1", "gcn.agg2", "gcn.gconv2",
                 "gcn.tconv2", "gcn.fcn")


class MMEgoError(RuntimeError):
    pass


class RawFramesStruct(C.Structure):
    """mmego_raw_frames_t of include/mmego_b200.h (device pointers)."""
    _fields_ = [(n, _vp) for n in ("points", "pt_start", "key", "imu", "R_btc", "t_R0R", "R_ref", "orientation_ref")] + \
               [("n_frames", _ll)]
    DTYPES = dict(points=torch.float32, pt_start=torch.int64, key=torch.float64, imu=torch.float64, R_btc=torch.float64,
                  t_R0R=torch.float64, R_ref=torch.float64, orientation_ref=torch.float64)


class Lib:
    """A loaded libmmego_b200 (or, in tests only, the emulated build of the same sources)."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise MMEgoError(
                f"{path} not found: build it with `python -m mmego_b200.build` (needs nvcc). "
                "mmego_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        from . import ABI_VERSION
        got = self.dll.mmego_abi_version()
        if got != ABI_VERSION:
            raise MMEgoError(f"{path}: ABI version {got}, expected {ABI_VERSION}")


_LIB: Optional[Lib] = None


def load() -> Lib:
    """The product library.  Raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        _LIB = Lib(LIB_PATH)
    return _LIB


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Handle:
    """One mmego_handle (one per GPU).  ``require_cuda=False`` exists for the emulator tests only."""

    def __init__(self, device: Optional[torch.device] = None, lib: Optional[Lib] = None, require_cuda: bool = True):
        self.lib = lib or load()
        self.require_cuda = require_cuda
        if require_cuda:
            if not torch.cuda.is_available():
                raise MMEgoError("mmego_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device(device if device is not None else "cuda")
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            self.device = device
            index = device.index
        else:
            self.device = torch.device("cpu")
            index = 0
        hp = _vp()
        rc = self.lib.dll.mmego_create(C.byref(hp), index)
        if rc != 0:
            raise MMEgoError(f"mmego_create failed ({rc}): {self.lib.dll.mmego_last_error(None).decode()}")
        self._h = hp
        self._ws: Optional[torch.Tensor] = None
        self._keep = []
        self.weights_owner: Dict[int, object] = {}     # net id -> (module id, tensor versions) of the packed set (engine.NativeNet)

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dll.mmego_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise MMEgoError(f"{what} failed ({rc}): {self.lib.dll.mmego_last_error(self._h).decode()}")

    def _t(self, t: Optional[torch.Tensor], name: str, dtype=torch.float32) -> Optional[torch.Tensor]:
        if t is None:
            return None
        if self.require_cuda and (not t.is_cuda or t.device != self.device):
            raise MMEgoError(f"{name} must live on {self.device} (got {t.device}); mmego_b200 has no CPU path")
        if t.dtype != dtype:
            raise MMEgoError(f"{name} must be {dtype} (got {t.dtype})")
        if not t.is_contiguous():
            raise MMEgoError(f"{name} must be contiguous")
        return t

    def _stream(self) -> Optional[int]:
        if not self.require_cuda:
            return None
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.0) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def set_option(self, key: str, value: int):
        self._ck(self.lib.dll.mmego_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def set_weights(self, net: int, state_dict: Mapping[str, torch.Tensor]):
        items = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in state_dict.items()
                 if torch.is_tensor(v) and v.is_floating_point()]
        n = len(items)
        names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (_vp * n)(*[v.data_ptr() for _, v in items])
        numels = (_ll * n)(*[v.numel() for _, v in items])
        self.weights_owner.pop(net, None)              # a direct upload invalidates any module's claim on this slot
        self._ck(self.lib.dll.mmego_set_weights(self._h, net, names, ptrs, numels, n), "set_weights")

    def workspace_bytes(self, stage: int, B: int, L: int, N: int, n_imu: int) -> int:
        return int(self.lib.dll.mmego_workspace_bytes(self._h, stage, B, L, N, n_imu))

    def launch_count(self) -> int:
        return int(self.lib.dll.mmego_launch_count(self._h))

    def profile_begin(self):
        self._ck(self.lib.dll.mmego_profile_begin(self._h), "profile_begin")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for name in PROFILE_SPANS:
            ms, n, k = C.c_double(), _ll(), _ll()
            self._ck(self.lib.dll.mmego_profile_read(self._h, name.encode(), C.byref(ms), C.byref(n), C.byref(k)),
                     "profile_read")
            if k.value:
                out[name] = dict(ms=ms.value, launches=int(n.value), spans=int(k.value))
        return out

    def debug_stats(self, reset=True):
        buf = (C.c_ulonglong * 8)()
        self._ck(self.lib.dll.mmego_debug_stats(self._h, buf, 1 if reset else 0), "debug_stats")
        return list(buf)

    def profile_end(self):
        self._ck(self.lib.dll.mmego_profile_end(self._h), "profile_end")

    def tap(self, name: str, dst: torch.Tensor):
        self._keep.append(dst)
        self._ck(self.lib.dll.mmego_debug_tap(self._h, name.encode(), dst.data_ptr(), dst.numel() * dst.element_size()),
                 "debug_tap")

    # ------------------------------------------------------------------------------------------ stages
    def imu_forward(self, imu: torch.Tensor):
        imu = self._t(imu, "imu")
        B, L, n, c = imu.shape
        if c != 15:
            raise MMEgoError(f"imu must be [B,L,n,15] (got {tuple(imu.shape)})")
        R = torch.empty(B, L, 3, 3, dtype=torch.float32, device=imu.device)
        t = torch.empty(B, L, 3, dtype=torch.float32, device=imu.device)
        nb = self.workspace_bytes(STAGE_IMU, B, L, 0, n)
        ws = self._workspace(nb)
        self._ck(self.lib.dll.mmego_imu_forward(self._h, imu.data_ptr(), R.data_ptr(), t.data_ptr(), B, L, n,
                                                ws.data_ptr(), ws.numel(), self._stream()), "imu_forward")
        return R, t

    def upper_forward(self, x, h0, c0, initial_body, R, t, body_index_mode=BODY_REF, b_offset=0, B_global=None,
                      want_q=True, want_weights=True, want_state=True):
        x = self._t(x, "x")
        B, L, N, D = x.shape
        if D != 6:
            raise MMEgoError(f"x must be [B,L,N,6] (got {tuple(x.shape)})")
        h0, c0 = self._t(h0, "h0_g"), self._t(c0, "c0_g")
        for nm, s in (("h0_g", h0), ("c0_g", c0)):
            if s is not None and tuple(s.shape) != (6, B, 64):
                raise MMEgoError(f"{nm} must be [6,{B},64] (got {tuple(s.shape)})")
        body, R, t = self._t(initial_body, "initial_body"), self._t(R, "R"), self._t(t, "t")
        Bg = B if B_global is None else B_global
        if body.shape[0] < Bg or tuple(body.shape[1:]) != (20, 3):
            raise MMEgoError(f"initial_body must be [{Bg},20,3] (got {tuple(body.shape)})")
        if R.numel() != B * L * 9 or t.numel() != B * L * 3:
            raise MMEgoError("R must be [B,L,3,3] and t [B,L,3]")
        dev = x.device
        l = torch.empty(B, L, 15, 3, dtype=torch.float32, device=dev)
        q = torch.empty(B, L, 14, 3, 3, dtype=torch.float32, device=dev) if want_q else None
        gw = torch.empty(B * L, N, 1, dtype=torch.float32, device=dev) if want_weights else None
        hn = torch.empty(6, B, 64, dtype=torch.float32, device=dev) if want_state else None
        cn = torch.empty(6, B, 64, dtype=torch.float32, device=dev) if want_state else None
        ws = self._workspace(self.workspace_bytes(STAGE_UPPER, B, L, N, 0))
        self._ck(self.lib.dll.mmego_upper_forward(self._h, x.data_ptr(), _ptr(h0), _ptr(c0), body.data_ptr(), R.data_ptr(),
                                                  t.data_ptr(), l.data_ptr(), _ptr(q), _ptr(gw), _ptr(hn), _ptr(cn), B, L, N,
                                                  body_index_mode, b_offset, Bg, ws.data_ptr(), ws.numel(),
                                                  self._stream()), "upper_forward")
        return l, q, gw, hn, cn

    def lower_forward(self, upper_l, x, initial_body, R, t, body_index_mode=BODY_REF, b_offset=0, B_global=None,
                      want_q=True):
        x = self._t(x, "x")
        B, L, N, D = x.shape
        if D != 6:
            raise MMEgoError(f"x must be [B,L,N,6] (got {tuple(x.shape)})")
        upper_l = self._t(upper_l, "upper_l")
        if upper_l.numel() != B * L * 45:
            raise MMEgoError(f"upper_l must be [B,L,15,3] (got {tuple(upper_l.shape)})")
        body, R, t = self._t(initial_body, "initial_body"), self._t(R, "R"), self._t(t, "t")
        Bg = B if B_global is None else B_global
        if body.shape[0] < Bg or tuple(body.shape[1:]) != (20, 3):
            raise MMEgoError(f"initial_body must be [{Bg},20,3] (got {tuple(body.shape)})")
        dev = x.device
        l = torch.empty(B, L, 8, 3, dtype=torch.float32, device=dev)
        q = torch.empty(B, L, 6, 3, 3, dtype=torch.float32, device=dev) if want_q else None
        ws = self._workspace(self.workspace_bytes(STAGE_LOWER, B, L, N, 0))
        self._ck(self.lib.dll.mmego_lower_forward(self._h, upper_l.data_ptr(), x.data_ptr(), body.data_ptr(), R.data_ptr(),
                                                  t.data_ptr(), l.data_ptr(), _ptr(q), B, L, N, body_index_mode, b_offset,
                                                  Bg, ws.data_ptr(), ws.numel(), self._stream()), "lower_forward")
        return l, q

    def build_snippets(self, raw: Mapping[str, torch.Tensor], starts: torch.Tensor, slot_src: Optional[torch.Tensor] = None,
                       seed: int = 0, L: int = 20, N: int = 128):
        """raw: the arrays of mmego_raw_frames_t as tensors on this handle's device; starts [B] int64 (first source
        frame of every snippet); slot_src [B,L,N] int32 or None.  Returns dict(data, imu, key, R, t)."""
        st = RawFramesStruct()
        for name, dt in RawFramesStruct.DTYPES.items():
            st_t = self._t(raw[name], name, dt)
            setattr(st, name, st_t.data_ptr())
        st.n_frames = int(raw["pt_start"].numel()) - 1
        starts = self._t(starts, "starts", torch.int64)
        B = starts.numel()
        if B and (int(starts.min()) < 0 or int(starts.max()) + L > st.n_frames):
            raise MMEgoError(f"starts must keep every {L}-frame window inside the {st.n_frames} raw frames")
        if slot_src is not None:
            slot_src = self._t(slot_src, "slot_src", torch.int32)
            if slot_src.numel() != B * L * N:
                raise MMEgoError(f"slot_src must be [{B},{L},{N}] (got {tuple(slot_src.shape)})")
        dev = starts.device
        out = dict(data=torch.empty(B, L, N, 6, dtype=torch.float32, device=dev),
                   imu=torch.empty(B, L, 20, 15, dtype=torch.float32, device=dev),
                   key=torch.empty(B, L, 21, 3, dtype=torch.float32, device=dev),
                   R=torch.empty(B, L, 3, 3, dtype=torch.float32, device=dev),
                   t=torch.empty(B, L, 3, dtype=torch.float32, device=dev))
        self._ck(self.lib.dll.mmego_build_snippets(self._h, C.addressof(st), starts.data_ptr(), _ptr(slot_src), seed,
                                                   out["data"].data_ptr(), out["imu"].data_ptr(), out["key"].data_ptr(),
                                                   out["R"].data_ptr(), out["t"].data_ptr(), B, L, N, self._stream()),
                 "build_snippets")
        return out

    def gcn_extract_feature(self, x: torch.Tensor):
        x = self._t(x, "x")
        B, Cc, T, V, M = x.shape
        if (Cc, V, M) != (3, 15, 1):
            raise MMEgoError(f"x must be [B,3,T,15,1] (got {tuple(x.shape)})")
        out = torch.empty(B, T, 15, 64, dtype=torch.float32, device=x.device)
        ws = self._workspace(self.workspace_bytes(STAGE_GCN, B, T, 0, 0))
        self._ck(self.lib.dll.mmego_gcn_extract_feature(self._h, x.data_ptr(), out.data_ptr(), B, T, ws.data_ptr(),
                                                        ws.numel(), self._stream()), "gcn_extract_feature")
        return out

    def transform2h_(self, points, R, t):
        points, R, t = self._t(points, "points"), self._t(R, "R"), self._t(t, "t")
        n, D = points.shape[-2], points.shape[-1]
        F = points.numel() // (n * D)
        self._ck(self.lib.dll.mmego_transform2h(self._h, points.data_ptr(), R.data_ptr(), t.data_ptr(), F, n, D,
                                                self._stream()), "transform2h")
        return points

    def transform2r(self, points, R, t):
        points, R, t = self._t(points, "points"), self._t(R, "R"), self._t(t, "t")
        n = points.shape[-2]
        F = points.numel() // (n * 3)
        out = torch.empty_like(points)
        self._ck(self.lib.dll.mmego_transform2r(self._h, points.data_ptr(), R.data_ptr(), t.data_ptr(), out.data_ptr(), F,
                                                n, self._stream()), "transform2r")
        return out

    def assemble_metrics(self, upper_l, lower_l, target=None, sums=None, want_pred=True):
        upper_l, lower_l = self._t(upper_l, "upper_l"), self._t(lower_l, "lower_l")
        B, L = upper_l.shape[:2]
        target = self._t(target, "target")
        sums = self._t(sums, "sums", torch.float64)
        if sums is not None and sums.numel() < SUMS_LEN:
            raise MMEgoError(f"sums must hold {SUMS_LEN} float64")
        pred = torch.empty(B, L, 21, 3, dtype=torch.float32, device=upper_l.device) if want_pred else None
        self._ck(self.lib.dll.mmego_assemble_metrics(self._h, upper_l.data_ptr(), lower_l.data_ptr(), _ptr(target),
                                                     _ptr(pred), _ptr(sums), B, L, self._stream()), "assemble_metrics")
        return pred

    def pipeline_forward(self, imu, x, initial_body, target=None, sums=None, body_index_mode=BODY_REF, b_offset=0,
                         B_global=None, outs: Optional[Dict[str, torch.Tensor]] = None, want_pred=True):
        imu, x = self._t(imu, "imu"), self._t(x, "x")
        B, L, N, _ = x.shape
        n = imu.shape[2]
        body = self._t(initial_body, "initial_body")
        Bg = B if B_global is None else B_global
        target = self._t(target, "target")
        sums = self._t(sums, "sums", torch.float64)
        self._check_batch(imu, x, body, target, sums, Bg, b_offset)
        dev = x.device
        pred = torch.empty(B, L, 21, 3, dtype=torch.float32, device=dev) if want_pred else None
        o = outs if outs is not None else {}
        ws = self._workspace(self.workspace_bytes(STAGE_PIPELINE, B, L, N, n))
        self._ck(self.lib.dll.mmego_pipeline_forward(
            self._h, imu.data_ptr(), x.data_ptr(), body.data_ptr(), _ptr(target), _ptr(pred), _ptr(sums),
            _ptr(o.get("R")), _ptr(o.get("t")), _ptr(o.get("upper_l")), _ptr(o.get("lower_l")), B, L, N, n, body_index_mode,
            b_offset, Bg, ws.data_ptr(), ws.numel(), self._stream()), "pipeline_forward")
        return pred

    @staticmethod
    def _check_batch(imu, x, body, target, sums, Bg, b_offset):
        """The C entry points trust the sizes they are given: everything they will read is checked here."""
        if x.dim() != 4 or x.shape[-1] != 6:
            raise MMEgoError(f"data must be [B,L,N,6] (got {tuple(x.shape)})")
        B, L = x.shape[:2]
        if imu.dim() != 4 or imu.shape[-1] != 15 or tuple(imu.shape[:2]) != (B, L):
            raise MMEgoError(f"imu must be [{B},{L},n_imu,15] to match data (got {tuple(imu.shape)})")
        if b_offset < 0 or Bg < b_offset + B:
            raise MMEgoError(f"bad shard: b_offset {b_offset} + B {B} exceeds B_global {Bg}")
        if body.dim() != 3 or body.shape[0] < Bg or tuple(body.shape[1:]) != (20, 3):
            raise MMEgoError(f"initial_body must be [{Bg},20,3] -- one skeleton per snippet of the GLOBAL batch "
                             f"(got {tuple(body.shape)})")
        if target is not None and target.numel() != B * L * 63:
            raise MMEgoError(f"target must be [{B},{L},21,3] (got {tuple(target.shape)})")
        if sums is not None and sums.numel() < SUMS_LEN:
            raise MMEgoError(f"sums must hold {SUMS_LEN} float64")

    def infer_host(self, imu, data, initial_body, target=None, body_index_mode=BODY_REF, b_offset=0, B_global=None,
                   out_pred=None, out_sums=None):
        """HOST tensors in, HOST tensors out (the call a reference-side binding makes per batch).  `out_pred`
        [B,L,21,3] float32 / `out_sums` [46] float64: caller-owned (ideally pinned) result buffers; allocated per call
        when omitted (a fresh pinned allocation costs ~10 ms per 20 MB)."""
        for nm, v in (("imu", imu), ("data", data), ("initial_body", initial_body), ("target", target)):
            if v is not None and (v.is_cuda or v.dtype != torch.float32 or not v.is_contiguous()):
                raise MMEgoError(f"{nm} must be a contiguous float32 HOST tensor")
        B, L, N, _ = data.shape
        n = imu.shape[2]
        Bg = B if B_global is None else B_global
        self._check_batch(imu, data, initial_body, target, None, Bg, b_offset)
        pin = self.require_cuda
        if out_pred is not None:
            if out_pred.is_cuda or out_pred.dtype != torch.float32 or not out_pred.is_contiguous() or out_pred.numel() != B * L * 63:
                raise MMEgoError(f"out_pred must be a contiguous float32 HOST tensor of {B * L * 63} elements")
            pred = out_pred
        else:
            pred = torch.empty(B, L, 21, 3, dtype=torch.float32, pin_memory=pin)
        sums = None
        if target is not None:
            if out_sums is not None:
                if out_sums.is_cuda or out_sums.dtype != torch.float64 or out_sums.numel() != SUMS_LEN:
                    raise MMEgoError(f"out_sums must be a float64 HOST tensor of {SUMS_LEN} elements")
                sums = out_sums
            else:
                sums = torch.zeros(SUMS_LEN, dtype=torch.float64, pin_memory=pin)
        self._ck(self.lib.dll.mmego_infer_host(self._h, imu.data_ptr(), data.data_ptr(), initial_body.data_ptr(),
                                               _ptr(target), pred.data_ptr(), _ptr(sums), B, L, N, n, body_index_mode,
                                               b_offset, Bg), "infer_host")
        return pred, sums
