"""MLP actor-critic (64-256-128-64 ELU, mean 12 + value 1) with the forward pass on the tcgen05 kernel.

Parameters are ordinary ``nn.Linear`` modules in the layout of the reference's ``Shared`` model
(scripts/skrl_ppo_locomanipulation.py:33-42) / rl_games ``actor_critic`` network
(cfg/train/JointLocomanipulationPPO.yaml:10-32), so a reference ``state_dict`` loads unchanged;
``forward`` packs them (bf16, UMMA canonical layout) and calls ``lm_mlp_forward``.  No eager fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib


class MlpActorCritic(nn.Module):
    def __init__(self, num_observations: int = 64, num_actions: int = 12, device="cuda:0"):
        super().__init__()
        assert num_observations == 64 and num_actions == 12, "kernel is specialised to the tasks' 64 -> 12+1 network"
        self.net = nn.Sequential(nn.Linear(64, 256), nn.ELU(), nn.Linear(256, 128), nn.ELU(), nn.Linear(128, 64), nn.ELU())
        self.mean_layer = nn.Linear(64, num_actions)
        self.log_std_parameter = nn.Parameter(torch.zeros(num_actions))
        self.value_layer = nn.Linear(64, 1)
        self.to(device)
        self._device = torch.device(device)
        self._lib = None
        self._handle = None
        self._packed_version = None

    # ---------------------------------------------------------------- kernel plumbing
    def _ensure(self):
        if self._handle is None:
            if self._device.type != "cuda" or not torch.cuda.is_available():
                raise _lib.LmError("MlpActorCritic.forward needs a CUDA device (no CPU fallback)")
            self._lib = _lib.load()
            idx = self._device.index if self._device.index is not None else torch.cuda.current_device()
            self._handle = self._lib.lm_mlp_create(idx)
            if not self._handle:
                raise _lib.LmError(f"lm_mlp_create failed: {self._lib.lm_last_error().decode()}")

    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def pack(self) -> None:
        """Convert the current fp32 parameters into the kernel's resident bf16 blob."""
        self._ensure()
        l1, l2, l3 = self.net[0], self.net[2], self.net[4]
        ps = [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, self.mean_layer.weight, self.mean_layer.bias,
              self.value_layer.weight, self.value_layer.bias]
        ts = [p.detach().float().contiguous() for p in ps]
        stream = torch.cuda.current_stream(self._device).cuda_stream
        _lib.check(self._lib.lm_mlp_pack(self._handle, *[t.data_ptr() for t in ts], stream), "lm_mlp_pack")
        self._keep = ts
        self._packed_version = self._param_version()

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, norm_mean: Optional[torch.Tensor] = None, norm_inv_std: Optional[torch.Tensor] = None,
                norm_clip: float = 5.0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """obs [N,64] fp32 -> (mu [N,12], value [N,1], log_std [12])."""
        self._ensure()
        if self._packed_version != self._param_version():
            self.pack()
        obs = obs.to(self._device, dtype=torch.float32).contiguous()
        n = obs.shape[0]
        if out is None:
            mu = torch.empty((n, 12), device=self._device, dtype=torch.float32)
            value = torch.empty((n, 1), device=self._device, dtype=torch.float32)
        else:
            mu, value = out
        stream = torch.cuda.current_stream(self._device).cuda_stream
        rc = self._lib.lm_mlp_forward(self._handle, obs.data_ptr(), _lib.ptr(norm_mean), _lib.ptr(norm_inv_std), float(norm_clip),
                                      mu.data_ptr(), value.data_ptr(), n, stream)
        _lib.check(rc, "lm_mlp_forward")
        return mu, value, self.log_std_parameter

    def reference_forward(self, obs: torch.Tensor):
        """The same network evaluated by stock torch ops (used by tests as the fp32 reference)."""
        h = self.net(obs)
        return self.mean_layer(h), self.value_layer(h), self.log_std_parameter

    def __del__(self):
        try:
            if self._handle:
                self._lib.lm_mlp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
