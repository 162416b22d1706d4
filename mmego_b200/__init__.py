"""mmego_b200 -- B200 (sm_100a) implementation of mmEgo's inference forward pass
(IMU_Net -> Upper_Net -> Lower_Net/ST-GCN -> joint decode) behind the reference's Python surface.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute is hand-written CUDA in
``csrc/`` reached through the C ABI of ``include/mmego_b200.h`` (``_capi.py``).  No CPU fallback.
"""
ABI_VERSION = 2

__all__ = ["ABI_VERSION"]
