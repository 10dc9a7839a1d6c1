"""GNN limb-graph actor-critic with the forward pass on the fused kernels: bf16 tcgen05 tensor cores
(csrc/lm_gnn_tc.cu, ``precision="bf16"``, the default: logits within 1e-2 of fp32) or fp32 CUDA cores
(csrc/lm_gnn.cu, ``precision="fp32"``: within 1e-5 of the reference).

Module / parameter names follow the reference (scripts/graph_model_orebot_ov.py:82-241 and the ``Shared``
wrapper at scripts/skrl_ppo_locomanipulation.py:44-49) so its checkpoints load unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib

HIDDEN = 32


class _GraphLayer(nn.Module):
    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.linear1 = nn.Linear(in_features * 2, hidden_features)
        self.linear2 = nn.Linear(hidden_features, out_features)


class _GraphNet(nn.Module):
    def __init__(self, hidden_features=HIDDEN, out_features=HIDDEN):
        super().__init__()
        self.input_layer1 = nn.Linear(16, hidden_features)
        self.input_layer2 = nn.Linear(4, hidden_features)
        self.graph_layer1 = _GraphLayer(hidden_features, hidden_features, hidden_features)
        self.graph_layer2 = _GraphLayer(hidden_features, hidden_features, hidden_features)
        self.graph_layer3 = _GraphLayer(hidden_features, hidden_features, out_features)


class _Head(nn.Module):
    def __init__(self, hidden_features):
        super().__init__()
        self.action_layer = nn.Linear(hidden_features, 1)


class GnnActorCritic(nn.Module):
    def __init__(self, num_actions: int = 12, device="cuda:0", precision: str = "bf16"):
        super().__init__()
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        self.net = _GraphNet()
        self.mean_layer = _Head(HIDDEN)
        self.log_std_parameter = nn.Parameter(torch.zeros(num_actions))
        self.value_layer = _Head(HIDDEN)
        self.to(device)
        self._device = torch.device(device)
        self._lib = None
        self._handle = None
        self._packed_version = None

    def _ensure(self):
        if self._handle is None:
            if self._device.type != "cuda" or not torch.cuda.is_available():
                raise _lib.LmError("GnnActorCritic.forward needs a CUDA device (no CPU fallback)")
            self._lib = _lib.load()
            idx = self._device.index if self._device.index is not None else torch.cuda.current_device()
            self._handle = self._lib.lm_gnn_create(idx)
            if not self._handle:
                raise _lib.LmError(f"lm_gnn_create failed: {self._lib.lm_last_error().decode()}")

    def ordered_parameters(self):
        n = self.net
        ps = [n.input_layer1.weight, n.input_layer1.bias, n.input_layer2.weight, n.input_layer2.bias]
        for g in (n.graph_layer1, n.graph_layer2, n.graph_layer3):
            ps += [g.linear1.weight, g.linear1.bias, g.linear2.weight, g.linear2.bias]
        ps += [self.mean_layer.action_layer.weight, self.mean_layer.action_layer.bias,
               self.value_layer.action_layer.weight, self.value_layer.action_layer.bias]
        return ps

    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())

    def mark_updated(self) -> None:
        """Call after writing the parameters behind torch's back (optimiser kernels): the next forward re-packs."""
        self._packed_version = None

    def pack(self) -> None:
        self._ensure()
        ts = [p.detach().float().contiguous() for p in self.ordered_parameters()]
        arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        stream = torch.cuda.current_stream(self._device).cuda_stream
        _lib.check(self._lib.lm_gnn_pack(self._handle, arr, len(ts), stream), "lm_gnn_pack")
        self._keep = ts
        self._packed_version = self._param_version()

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
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
        fn = self._lib.lm_gnn_forward_bf16 if self.precision == "bf16" else self._lib.lm_gnn_forward
        with _lib.nvtx("gnn_forward"):
            rc = fn(self._handle, obs.data_ptr(), mu.data_ptr(), value.data_ptr(), n, stream)
        _lib.check(rc, "lm_gnn_forward")
        return mu, value, self.log_std_parameter

    def reference_forward(self, obs: torch.Tensor):
        """The same network in stock torch ops, differentiable: the training path (autograd) and the tests' fp32 reference.
        Per-endpoint form of graph_model_orebot_ov.py:46-59,142-159 — P = W1[:, :32] h, Q = W1[:, 32:] h per node, then per edge
        ELU(W2 ELU(P_target + Q_source + b1) + b2) and a max over the in-edges of each target."""
        F = torch.nn.functional
        n = obs.shape[0]
        net = self.net
        order = (0, 1, 2, 3, 4, 6, 8, 10, 5, 7, 9, 11)                          # :117-126
        cols = torch.tensor([[16 + j, 28 + j, 40 + j, 52 + j] for j in order], device=obs.device)
        h = torch.cat([net.input_layer1(obs[:, 0:16]).view(n, 1, -1), net.input_layer2(obs[:, cols])], dim=1)        # [n,13,32]
        src = [0] * 4 + list(range(1, 5)) + list(range(5, 9))
        dst = list(range(1, 5)) + list(range(5, 9)) + list(range(9, 13))
        s_idx = torch.tensor(src + dst, device=obs.device)                       # both directions (:142-159)
        t_idx = torch.tensor(dst + src, device=obs.device)
        for layer in (net.graph_layer1, net.graph_layer2, net.graph_layer3):
            w1, hd = layer.linear1.weight, h.shape[-1]
            pq = F.linear(h, torch.cat([w1[:, :hd], w1[:, hd:]], dim=0))         # [n,13,64]: P | Q
            m = F.elu(pq[:, t_idx, :hd] + pq[:, s_idx, hd:] + layer.linear1.bias)
            m = F.elu(layer.linear2(m))                                          # [n,24,32]
            out = torch.full((n, 13, m.shape[-1]), -float("inf"), device=obs.device, dtype=m.dtype)
            h = out.scatter_reduce(1, t_idx.view(1, -1, 1).expand_as(m), m, reduce="amax", include_self=True)
        mu = self.mean_layer.action_layer(h[:, 1:13]).squeeze(-1)
        value = self.value_layer.action_layer(torch.max(h, dim=1).values)
        return mu, value, self.log_std_parameter

    def __del__(self):
        try:
            if self._handle:
                self._lib.lm_gnn_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
