"""Rollout buffers and running mean/std on the GPU (csrc/lm_rollout.cu): the caller-side formats around the task/policy path.

Names follow rl_games 1.6.0 (``ExperienceBuffer.tensor_dict`` keys, ``RunningMeanStd`` buffers ``running_mean`` /
``running_var`` / ``count``) so that its PPO update can read the tensors unchanged; ``variant="skrl"`` switches the
normaliser to skrl's ``RunningStandardScaler`` form.  Buffers are time-major ``[T, N, ...]``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib

RLGAMES, SKRL = 0, 1


def _handle(device: torch.device):
    if device.type != "cuda" or not torch.cuda.is_available():
        raise _lib.LmError("rollout kernels need a CUDA device (no CPU fallback)")
    lib = _lib.load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = lib.lm_rollout_create(idx)
    if not h:
        raise _lib.LmError(f"lm_rollout_create failed: {lib.lm_last_error().decode()}")
    return lib, h


class RunningMeanStd:
    """float64 running mean / variance / count, updated by ``lm_rms_update`` (or fused into ``RolloutBuffer.write``)."""

    def __init__(self, dim: int, device="cuda:0", variant: int = RLGAMES):
        self.device = torch.device(device)
        self.dim, self.variant = int(dim), int(variant)
        self.running_mean = torch.zeros(dim, dtype=torch.float64, device=self.device)
        self.running_var = torch.ones(dim, dtype=torch.float64, device=self.device)
        self.count = torch.ones((), dtype=torch.float64, device=self.device)
        self._mean32 = torch.zeros(dim, dtype=torch.float32, device=self.device)
        self._inv_std32 = torch.ones(dim, dtype=torch.float32, device=self.device)
        self._lib, self._h = None, None

    def c_struct(self) -> _lib.LmRunningStats:
        return _lib.LmRunningStats(self.running_mean.data_ptr(), self.running_var.data_ptr(), self.count.data_ptr(), self.dim, self.variant)

    def update(self, x: torch.Tensor) -> None:
        if self._h is None:
            self._lib, self._h = _handle(self.device)
        x = x.to(self.device, dtype=torch.float32).contiguous()
        n = x.numel() // self.dim
        st = self.c_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.lm_rms_update(self._h, x.data_ptr(), n, C.byref(st), stream), "lm_rms_update")

    def export(self):
        """(mean, inv_std) fp32 vectors in the form ``lm_mlp_forward`` takes."""
        lib = _lib.load()
        st = self.c_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.lm_rms_export(C.byref(st), self._mean32.data_ptr(), self._inv_std32.data_ptr(), stream), "lm_rms_export")
        return self._mean32, self._inv_std32

    def export_denorm(self) -> torch.Tensor:
        """[2] = (mean, std) fp32 of a scalar normaliser, the form ``RolloutBuffer.write(value_denorm=...)`` takes."""
        if getattr(self, "_denorm32", None) is None:
            self._denorm32 = torch.zeros(2, dtype=torch.float32, device=self.device)
        st = self.c_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().lm_rms_export_denorm(C.byref(st), self._denorm32.data_ptr(), stream), "lm_rms_export_denorm")
        return self._denorm32

    def normalize(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, clip: float = 5.0) -> torch.Tensor:
        """Scalar normaliser only: clamp((x - mean) / std, +-clip) with the CURRENT statistics (eval-mode forward)."""
        assert self.dim == 1
        mean32, inv_std32 = self.export()
        out = torch.empty_like(x) if out is None else out
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().lm_normalize_clamp(x.data_ptr(), out.data_ptr(), x.numel(), mean32.data_ptr(), inv_std32.data_ptr(), float(clip), stream),
                   "lm_normalize_clamp")
        return out

    def __del__(self):
        try:
            if self._h:
                self._lib.lm_rollout_destroy(self._h)
                self._h = None
        except Exception:
            pass


class RolloutBuffer:
    """Time-major rollout storage written by ONE kernel launch per env step."""

    def __init__(self, horizon: int, num_envs: int, device="cuda:0", reward_scale: float = 1.0):
        self.device = torch.device(device)
        self.horizon, self.num_envs, self.reward_scale = int(horizon), int(num_envs), float(reward_scale)
        T, N, d = self.horizon, self.num_envs, self.device
        self.tensor_dict: Dict[str, torch.Tensor] = {
            "obses": torch.zeros(T, N, 64, device=d), "actions": torch.zeros(T, N, 12, device=d), "mus": torch.zeros(T, N, 12, device=d),
            "values": torch.zeros(T, N, device=d), "rewards": torch.zeros(T, N, device=d), "dones": torch.zeros(T, N, device=d),
            "neglogpacs": torch.zeros(T, N, device=d)}
        self._lib, self._h = _handle(self.device)
        self._sample_calls = 0                    # stream position of the kernel's own generator
        td = self.tensor_dict
        self._buf = _lib.LmRolloutBuffers(td["obses"].data_ptr(), td["actions"].data_ptr(), td["mus"].data_ptr(), td["values"].data_ptr(),
                                          td["rewards"].data_ptr(), td["dones"].data_ptr(), td["neglogpacs"].data_ptr(), T, N, self.reward_scale)

    def _step(self, obs, mu, value, log_std, rew, dones, noise, sample_seed, time_outs, gamma, value_denorm, sink=None):
        in_kernel = noise is None and sample_seed is not None
        step = _lib.LmRolloutStep(_lib.ptr(obs), _lib.ptr(mu), _lib.ptr(value), _lib.ptr(log_std), _lib.ptr(noise), _lib.ptr(rew), _lib.ptr(dones),
                                  int(sample_seed or 0), self._sample_calls, int(in_kernel), 0, _lib.ptr(time_outs), float(gamma), 0,
                                  _lib.ptr(value_denorm), C.addressof(sink) if sink is not None else None)
        self._keep_sink = sink
        return step, in_kernel

    def write(self, t: int, obs, mu, value, log_std, rew, dones, noise: Optional[torch.Tensor] = None,
              actions_out: Optional[torch.Tensor] = None, obs_stats: Optional[RunningMeanStd] = None, sample_seed: Optional[int] = None,
              time_outs: Optional[torch.Tensor] = None, gamma: float = 0.99, value_denorm: Optional[torch.Tensor] = None, sink=None) -> None:
        """``sink`` (``LocoManipulationTask.action_sink()``): the sampled actions also go straight into the task's buffers (clamp,
        stash, take_action), and the next ``pre_physics_step`` is called with ``actions=None``.
        Slot ``t`` <- this step, in rl_games' ``play_steps`` order: ``obs`` / ``dones`` / ``rew`` are what the env returned last, so
        ``rew`` (the reward of the action stored in slot t-1) goes to slot t-1 — nothing at t = 0 — and :meth:`finish` closes the
        rollout.  ``noise`` = standard-normal draws [N,12] of the caller's generator; without it, ``sample_seed`` makes the kernel
        draw its own N(0,1) noise (counter-based: reproducible for a given seed and call count, no draw tensor), and neither means
        act = mu (evaluation).  ``actions_out`` receives the sampled actions for the next ``pre_physics_step``.  ``time_outs``
        (int64 [N]) switches on value_bootstrap; ``value_denorm`` ([2] = mean, std from ``RunningMeanStd.export_denorm``) stores
        de-normalised values (normalize_value)."""
        step, in_kernel = self._step(obs, mu, value, log_std, rew, dones, noise, sample_seed, time_outs, gamma, value_denorm, sink)
        if in_kernel:
            self._sample_calls += 1
        st = obs_stats.c_struct() if obs_stats is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.lm_rollout_write(self._h, C.byref(step), C.byref(self._buf), int(t), _lib.ptr(actions_out),
                                        C.byref(st) if st is not None else None, stream)
        _lib.check(rc, "lm_rollout_write")

    def finish(self, rew, value, dones, time_outs: Optional[torch.Tensor] = None, gamma: float = 0.99,
               value_denorm: Optional[torch.Tensor] = None):
        """The env step after slot T-1: its reward completes slot T-1; returns (last_values, last_dones) [N], GAE's bootstrap."""
        if not hasattr(self, "last_values"):
            self.last_values = torch.empty(self.num_envs, device=self.device)
            self.last_dones = torch.empty(self.num_envs, device=self.device)
        step, _ = self._step(None, None, value, None, rew, dones, None, None, time_outs, gamma, value_denorm)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.lm_rollout_finish(self._h, C.byref(step), C.byref(self._buf), self.last_values.data_ptr(), self.last_dones.data_ptr(), stream)
        _lib.check(rc, "lm_rollout_finish")
        return self.last_values, self.last_dones

    def __del__(self):
        try:
            if self._h:
                self._lib.lm_rollout_destroy(self._h)
                self._h = None
        except Exception:
            pass
