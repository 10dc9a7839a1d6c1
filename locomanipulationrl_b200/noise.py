"""Observation / action noise of the vec-env step on the GPU (csrc/lm_noise.cu).

Mirrors the reference ``Randomizer`` (utils/domain_randomization/randomize.py:156-305): the same yaml dictionaries
(``on_reset`` / ``on_interval`` with ``operation``, ``distribution``, ``distribution_parameters``, ``frequency_interval``),
the same buffers (``_observations_counter_buffer`` ...) and the two methods ``VecEnvRLGames.step`` calls
(envs/vec_env_rlgames.py:57-58,70-72).  No host synchronisation: the reference's ``nonzero()`` compaction is replaced by
per-env flags, and the draws come from the kernel's own counter-based generator unless the caller supplies them.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib

_OPS = {"additive": 0, "scaling": 1}
_DISTS = {"gaussian": 0, "normal": 0, "uniform": 1, "loguniform": 2, "log_uniform": 2}


def _spec(params: Optional[dict], interval: bool) -> _lib.LmNoiseSpec:
    if params is None:
        return _lib.LmNoiseSpec(0, 0, 0, 0, 0.0, 0.0)
    need = {"operation", "distribution", "distribution_parameters"} | ({"frequency_interval"} if interval else set())
    if not need.issubset(params.keys()):
        raise ValueError(f"Please ensure the following randomization parameters are provided: {sorted(need)}")     # randomize.py:162-164
    if params["operation"] not in _OPS or params["distribution"] not in _DISTS:
        raise ValueError(f"unsupported operation / distribution: {params['operation']} / {params['distribution']}")
    p0, p1 = (float(v) for v in params["distribution_parameters"])
    dist = _DISTS[params["distribution"]]
    if dist == 0:
        scale, shift = p1, p0                                   # torch.normal(mean, std): randn * std + mean
    elif dist == 1:
        scale, shift = p1 - p0, p0                              # (hi - lo) * rand + lo
    else:
        scale, shift = math.log(p1) - math.log(p0), math.log(p0)
    return _lib.LmNoiseSpec(1, _OPS[params["operation"]], dist, int(params.get("frequency_interval", 0)), scale, shift)


class NoiseRandomizer:
    """One buffer type (observations or actions)."""

    def __init__(self, params: dict, num_envs: int, dim: int, device="cuda:0", seed: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.LmError("NoiseRandomizer needs a CUDA device (no CPU fallback)")
        self.num_envs, self.dim, self.seed = int(num_envs), int(dim), int(seed)
        self.counter = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        self.correlated_noise = torch.zeros(num_envs, dim, device=self.device)
        self._on_reset, self._on_interval = _spec(params.get("on_reset"), False), _spec(params.get("on_interval"), True)
        self._lib = _lib.load()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = self._lib.lm_noise_create(idx)
        if not self._h:
            raise _lib.LmError(f"lm_noise_create failed: {self._lib.lm_last_error().decode()}")
        self._step = 0

    def apply(self, buffer: torch.Tensor, reset_buf: torch.Tensor, draws_reset=None, draws_interval=None, compact_draws: bool = False) -> torch.Tensor:
        """In place on ``buffer`` [N, dim] (fp32, contiguous); ``reset_buf`` int64 [N]."""
        st = _lib.LmNoiseState(self.counter.data_ptr(), self.correlated_noise.data_ptr(), self.num_envs, self.dim, self._on_reset, self._on_interval)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.lm_noise_apply(self._h, C.byref(st), _lib.ptr(buffer), _lib.ptr(reset_buf), _lib.ptr(draws_reset), _lib.ptr(draws_interval),
                                      int(compact_draws), self.seed, self._step, stream)
        _lib.check(rc, "lm_noise_apply")
        self._step += 1
        return buffer

    def __del__(self):
        try:
            if self._h:
                self._lib.lm_noise_destroy(self._h)
                self._h = None
        except Exception:
            pass


class Randomizer:
    """The two calls of ``VecEnvRLGames.step``, with the reference's names and yaml format."""

    def __init__(self, randomization_params: Dict[str, dict], num_envs: int, num_observations: int = 64, num_actions: int = 12,
                 device="cuda:0", seed: int = 0):
        self._obs = NoiseRandomizer(randomization_params["observations"], num_envs, num_observations, device, seed) \
            if randomization_params.get("observations") else None
        self._act = NoiseRandomizer(randomization_params["actions"], num_envs, num_actions, device, seed + 1) \
            if randomization_params.get("actions") else None
        self.randomize_observations, self.randomize_actions = self._obs is not None, self._act is not None

    def apply_observations_randomization(self, observations, reset_buf):
        return self._obs.apply(observations, reset_buf) if self._obs else observations

    def apply_actions_randomization(self, actions, reset_buf):
        return self._act.apply(actions, reset_buf) if self._act else actions
