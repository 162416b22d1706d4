"""Per-device library handles and the nn.Module base shared by the drop-in networks."""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch
from torch import nn

from . import _capi
from ._capi import MMEgoError

_HANDLES: Dict[int, "_capi.Handle"] = {}


def get_handle(device) -> "_capi.Handle":
    """One mmego_handle per GPU, created on first use.  Raises when CUDA or the library is unavailable."""
    device = torch.device(device)
    if device.type != "cuda":
        raise MMEgoError(f"mmego_b200 runs on CUDA devices only (got '{device}'); there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _HANDLES.get(idx)
    if h is None:
        h = _capi.Handle(torch.device("cuda", idx))
        _HANDLES[idx] = h
    return h


class NativeNet(nn.Module):
    """Parameters live in ordinary nn.Parameters/buffers (checkpoint-compatible); compute lives in the library.
    The packed device copies owned by the handle are refreshed whenever the module's tensors change."""

    _net_id: int = -1

    def __init__(self):
        super().__init__()

    def _weights_key(self):
        """(identity, version, storage) of every parameter and buffer, in state_dict order.  Walks the module tree directly:
        `state_dict()` builds prefixed names and runs its hooks on every call (220 us for LowerNet's 133 tensors -- a fifth
        of a batch-1 evaluation step on the host); the walk sees the same tensors in 1/3 of the time and still notices
        replaced Parameters, in-place edits (`_version`) and moved storage (`.to()`)."""
        out = []

        def walk(mod):
            for p in mod._parameters.values():
                if p is not None:
                    out.append((id(p), p._version, p.data_ptr()))
            for name, b in mod._buffers.items():
                if b is not None and name not in mod._non_persistent_buffers_set:
                    out.append((id(b), b._version, b.data_ptr()))
            for c in mod._modules.values():
                if c is not None:
                    walk(c)

        walk(self)
        return tuple(out)

    def _sync(self, device) -> "_capi.Handle":
        """A handle holds ONE packed weight set per net type and all modules on a GPU share the handle, so the handle
        records which module (and which version of its tensors) each slot currently holds: two UpperNet instances used
        alternately re-upload on every switch instead of silently running with each other's weights."""
        h = get_handle(device)
        key = (id(self), self._weights_key())
        if h.weights_owner.get(self._net_id) != key:
            h.set_weights(self._net_id, self.state_dict())
            h.weights_owner[self._net_id] = key
        return h

    @staticmethod
    def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
        if not torch.is_tensor(t):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise MMEgoError(f"{name} is on '{t.device}': mmego_b200 has no CPU path, move the batch to the B200")
        if t.dtype != torch.float32:
            raise MMEgoError(f"{name} must be float32 (got {t.dtype})")
        return t

    # -- same persistence surface as the reference classes (Net/IMU_Net.py:96-114 etc.)
    def save(self, name=None):
        if name is None:
            name = time.strftime("checkpoints/" + "%m%d_%H_%M_%S.pth")
        torch.save(self.state_dict(), name)
        return name

    def load(self, pathname):
        self.load_state_dict(torch.load(pathname, map_location="cpu", weights_only=True))
