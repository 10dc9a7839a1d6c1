"""PPO GAE on the GPU (csrc/lm_gae.cu): rl_games 1.6.0 ``discount_values`` and skrl ``compute_gae`` semantics."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

RLGAMES, SKRL = 0, 1


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, last_values: torch.Tensor,
                last_dones: Optional[torch.Tensor] = None, gamma: float = 0.99, lam: float = 0.95, variant: int = RLGAMES,
                env_major: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """rewards/values/dones: [T,N] (or [N,T] with ``env_major``) fp32 on a CUDA device; returns (advantages, returns)."""
    if not rewards.is_cuda:
        raise _lib.LmError("compute_gae runs on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    r = rewards.float().contiguous()
    v = values.float().contiguous().view_as(r)
    d = dones.float().contiguous().view_as(r)
    T, N = (r.shape[1], r.shape[0]) if env_major else (r.shape[0], r.shape[1])
    lv = last_values.float().contiguous().view(-1)
    ld = last_dones.float().contiguous().view(-1) if last_dones is not None else None
    adv = torch.empty_like(r)
    ret = torch.empty_like(r)
    stream = torch.cuda.current_stream(r.device).cuda_stream
    fn = lib.lm_gae_env_major if env_major else lib.lm_gae
    rc = fn(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), _lib.ptr(ld), adv.data_ptr(), ret.data_ptr(), T, N,
            float(gamma), float(lam), int(variant), stream)
    _lib.check(rc, "lm_gae")
    return adv, ret
