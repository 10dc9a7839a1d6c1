"""Host-buffer stepping of one GPU's env batch: the e2e entry a vec-env adapter with HOST tensors would call.

The batch is split into ``shards`` env shards — the same block sharding the multi-GPU path uses (envs
are independent; every shard keeps its own locomotion/manipulation halves) — each with its own task handle, CUDA
stream and pinned host buffers.  One ``step`` enqueues, per shard and without any host synchronisation in between,

    H2D actions + goal uniforms -> pre_physics_step -> H2D physics state -> post_physics_step -> D2H [obs/]rew/reset
    -> policy forward -> D2H mu/value                         (ONE library call: lm_task_step_host_policy_async)

(``readback_obs=False`` keeps the observations on the device for the device-resident policy, as the reference's adapter
does with ``rl_device = cuda``; the actions, value, reward and reset flags still return to the host every step)

then waits for every shard.  The physics-state copies of all shards go through ONE dedicated copy stream
(``LmHostStep.h2d_stream``): they queue back to back at full PCIe bandwidth — copies issued on different streams would run
concurrently and share the link, so every shard's state would land late — and shard k's kernels and device->host copies
run under shard k+1's state copy (PCIe is full duplex).  The step is then one pass of the state over PCIe plus the last
shard's kernels and read-back.  Inputs and outputs live in pinned host memory; nothing is cached across steps.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib, synthetic
from .configs import TaskConfig
from .tasks.base import LocoManipulationTask


class _Shard:
    def __init__(self, cfg: TaskConfig, n: int, device, seed: int, env_id_offset: int, readback_obs: bool = True):
        pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()
        self.n = n
        self.scene = synthetic.SyntheticScene(cfg, n, device=device, seed=seed, env_id_offset=env_id_offset)
        self.task = LocoManipulationTask(cfg, n, scene=self.scene, device=device, fuse_obs_clamp=True, seed=seed)
        self.stream = torch.cuda.Stream(device=device)
        self.actions_dev = torch.zeros(n, 12, device=device)
        self.uniforms_dev = torch.zeros(n, 3, device=device)
        self.mu_dev = torch.zeros(n, 12, device=device)
        self.value_dev = torch.zeros(n, 1, device=device)
        self.uniforms_host = pin(n, 3)
        self.obs_host, self.rew_host, self.reset_host = pin(n, 64), pin(n), pin(n, dtype=torch.long)
        self.mu_host = [pin(n, 12), pin(n, 12)]           # ping-pong: step t's mean is step t+1's action input
        self.value_host = pin(n, 1)
        self.io = _lib.LmHostStep()
        self.io.actions_dev, self.io.uniforms_dev = self.actions_dev.data_ptr(), self.uniforms_dev.data_ptr()
        self.io.uniforms_host = self.uniforms_host.data_ptr()
        # obs_host = NULL: the observations stay on the device for the device-resident policy (the reference's own data flow:
        # VecEnvRLGames hands them to rl_device = cuda); actions, value, reward and reset flags still return every step
        self.io.obs_host = self.obs_host.data_ptr() if readback_obs else None
        self.io.rew_host, self.io.reset_host = self.rew_host.data_ptr(), self.reset_host.data_ptr()
        self.pol = _lib.LmHostPolicy()
        self.pol.mu_dev, self.pol.value_dev = self.mu_dev.data_ptr(), self.value_dev.data_ptr()
        self.pol.value_host = self.value_host.data_ptr()


class ShardedHostStepper:
    def __init__(self, cfg: TaskConfig, num_envs: int, policy, shards: int = 4, device="cuda:0", seed: int = 42,
                 env_id_offset: int = 0, norm_mean: Optional[torch.Tensor] = None, norm_inv_std: Optional[torch.Tensor] = None,
                 readback_obs: bool = True, copy_stream: bool = True, shard_sizes: Optional[List[int]] = None):
        """``shard_sizes``: unequal shards (each an even number of envs, summing to ``num_envs``).  The step ends with the LAST shard's
        kernels and read-back, which nothing overlaps: a small last shard shortens that tail."""
        if shard_sizes is None:
            if num_envs % shards or (num_envs // shards) % 2:
                raise _lib.LmError("num_envs must split into an even number of envs per shard")
            shard_sizes = [num_envs // shards] * shards
        if sum(shard_sizes) != num_envs or any(s <= 0 or s % 2 for s in shard_sizes):
            raise _lib.LmError("shard_sizes must be positive even numbers summing to num_envs")
        shards = len(shard_sizes)
        self.lib = _lib.load()
        self.cfg, self.policy, self.device = cfg, policy, torch.device(device)
        self.norm_mean, self.norm_inv_std = norm_mean, norm_inv_std
        self.readback_obs = bool(readback_obs)
        offs = [sum(shard_sizes[:k]) for k in range(shards)]
        self.shards: List[_Shard] = [_Shard(cfg, shard_sizes[k], self.device, seed + 17 * k, env_id_offset + offs[k], self.readback_obs)
                                     for k in range(shards)]
        self.parity = 0
        # one copy stream for every shard's physics state (see the module docstring); a single shard needs none
        self.h2d_stream = torch.cuda.Stream(device=self.device) if (copy_stream and shards > 1) else None
        from .policy.mlp import MlpActorCritic
        policy._ensure()
        if getattr(policy, "_packed_version", None) != policy._param_version():
            policy.pack()
        kind = 1 if isinstance(policy, MlpActorCritic) else (2 if getattr(policy, "precision", "bf16") == "bf16" else 3)
        for s in self.shards:
            s.io.h2d_stream = self.h2d_stream.cuda_stream if self.h2d_stream is not None else None
            s.pol.kind, s.pol.handle = kind, policy._handle
            s.pol.norm_mean, s.pol.norm_inv_std, s.pol.norm_clip = _lib.ptr(norm_mean), _lib.ptr(norm_inv_std), 5.0
        self.h2d_bytes = sum(4 * (s.actions_dev.numel() + s.uniforms_dev.numel()) for s in self.shards)
        self.d2h_bytes = sum(4 * ((s.obs_host.numel() if self.readback_obs else 0) + s.rew_host.numel() + 12 * s.n + s.n) + 8 * s.reset_host.numel()
                             for s in self.shards)

    def step(self, host_states: List[Dict[str, torch.Tensor]]) -> None:
        """One env step for the whole batch.  ``host_states[k]`` = shard k's PhysX state in pinned host memory, packed like
        the shard's device arena (SyntheticScene.pack_host); goal uniforms are read from ``shards[k].uniforms_host``;
        actions are the previous step's policy means (``shards[k].mu_host``), as a host-side RL loop would feed back."""
        p = self.parity
        for s, hs in zip(self.shards, host_states):
            io, task = s.io, s.task
            io.actions_host = s.mu_host[p].data_ptr()
            for grp in ("loco", "mani"):
                if grp in hs:
                    setattr(io, grp + "_state_host", hs[grp].data_ptr())
                    setattr(io, grp + "_state_floats", hs[grp].numel())
                    setattr(io, grp + "_state_dev", s.scene.arena[grp].data_ptr())
            s.pol.mu_host = s.mu_host[p ^ 1].data_ptr()
            rc = self.lib.lm_task_step_host_policy_async(task._handle, task._lp(), task._mp(), task._env_pos.data_ptr(),
                                                         C.byref(task._buffers), C.byref(io), C.byref(s.pol), s.stream.cuda_stream)
            _lib.check(rc, "lm_task_step_host_policy_async")
            task._posted_since_pre = True
        for s in self.shards:
            _lib.check(self.lib.lm_stream_synchronize(s.stream.cuda_stream), "lm_stream_synchronize")
        self.parity = p ^ 1

    @property
    def mu_host(self) -> List[torch.Tensor]:
        return [s.mu_host[self.parity] for s in self.shards]
