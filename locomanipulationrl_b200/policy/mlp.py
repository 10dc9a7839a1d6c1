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


class _MlpTrainFunction(torch.autograd.Function):
    """Differentiable forward for callers that run their own loss through torch autograd (rl_games ``calc_gradients``, skrl's PPO
    ``_update``): the tcgen05 training forward saves the bf16 operand of every layer, ``backward`` runs ``lm_mlp_backward`` +
    ``lm_mlp_wgrad`` on them.  No gradient flows to the observations (they are data)."""

    @staticmethod
    def forward(ctx, model, obs, norm_mean, norm_inv_std, norm_clip, *params):
        lib, n, dev = model._lib, obs.shape[0], model._device
        ws = model._train_workspace(n)
        mu = torch.empty((n, 12), device=dev, dtype=torch.float32)
        value = torch.empty((n, 1), device=dev, dtype=torch.float32)
        s = torch.cuda.current_stream(dev).cuda_stream
        p = _lib.ptr
        _lib.check(lib.lm_mlp_forward_train(model._handle, p(obs), p(norm_mean), p(norm_inv_std), float(norm_clip), p(mu), p(value),
                                            p(ws["a0"]), p(ws["a1"]), p(ws["a2"]), p(ws["a3"]), n, s), "lm_mlp_forward_train")
        ctx.model, ctx.n, ctx.ws = model, n, ws
        return mu, value

    @staticmethod
    def backward(ctx, grad_mu, grad_value):
        model, n, ws = ctx.model, ctx.n, ctx.ws
        lib, dev = model._lib, model._device
        s = torch.cuda.current_stream(dev).cuda_stream
        p = _lib.ptr
        gm = (grad_mu if grad_mu is not None else torch.zeros(n, 12, device=dev)).float().contiguous()
        gv = (grad_value if grad_value is not None else torch.zeros(n, 1, device=dev)).float().contiguous().view(-1)
        _lib.check(lib.lm_mlp_backward(model._handle, p(gm), p(gv), p(ws["a1"]), p(ws["a2"]), p(ws["a3"]), p(ws["dh"]), p(ws["dz3"]), p(ws["dz2"]),
                                       p(ws["dz1"]), n, s), "lm_mlp_backward")
        grads = [torch.empty_like(q) for q in model._kernel_params()]
        _lib.check(lib.lm_mlp_wgrad(p(ws["dz1"]), p(ws["dz2"]), p(ws["dz3"]), p(ws["dh"]), p(ws["a0"]), p(ws["a1"]), p(ws["a2"]), p(ws["a3"]), n,
                                    p(ws["partials"]), ws["ctas"], *[g.data_ptr() for g in grads], s), "lm_mlp_wgrad")
        return (None, None, None, None, None, *grads)


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

    def mark_updated(self) -> None:
        """Call after writing the parameters behind torch's back (kernels writing through ``data_ptr``: torch's version counters do
        not move): the next forward re-packs the bf16 operand blobs."""
        self._packed_version = None

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

    def _kernel_params(self):
        l1, l2, l3 = self.net[0], self.net[2], self.net[4]
        return [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, self.mean_layer.weight, self.mean_layer.bias,
                self.value_layer.weight, self.value_layer.bias]

    def _train_workspace(self, n: int):
        """Activation / gradient tile blobs of the differentiable forward, grown on demand and reused between calls."""
        ws = getattr(self, "_train_ws", None)
        if ws is None or ws["rows"] < n:
            lib, dev = self._lib, self._device
            nb = lambda d: torch.empty(lib.lm_mlp_blob_bytes(n, d), dtype=torch.uint8, device=dev)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            ctas = lib.lm_device_sm_count(idx)
            ws = dict(rows=n, a0=nb(64), a1=nb(256), a2=nb(128), a3=nb(64), dz1=nb(256), dz2=nb(128), dz3=nb(64), dh=nb(16), ctas=ctas,
                      partials=torch.empty(ctas * lib.lm_mlp_wgrad_workspace_floats(), device=dev))
            self._train_ws = ws
        return ws

    def forward(self, obs: torch.Tensor, norm_mean: Optional[torch.Tensor] = None, norm_inv_std: Optional[torch.Tensor] = None,
                norm_clip: float = 5.0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, sink=None):
        """obs [N,64] fp32 -> (mu [N,12], value [N,1], log_std [12]).  Under ``torch.enable_grad()`` with trainable parameters the
        outputs carry an autograd graph whose backward runs on the tcgen05 kernels (one saved-activation workspace per module: take
        the backward of a forward before calling the next differentiable forward); under ``no_grad`` it is the inference kernel."""
        self._ensure()
        if self._packed_version != self._param_version():
            self.pack()
        obs = obs.to(self._device, dtype=torch.float32).contiguous()
        if torch.is_grad_enabled() and out is None and any(q.requires_grad for q in self._kernel_params()):
            mu, value = _MlpTrainFunction.apply(self, obs.detach(), norm_mean, norm_inv_std, norm_clip, *self._kernel_params())
            return mu, value, self.log_std_parameter
        return self._forward_inference(obs, norm_mean, norm_inv_std, norm_clip, out, sink)

    @torch.no_grad()
    def _forward_inference(self, obs, norm_mean, norm_inv_std, norm_clip, out, sink=None):
        """``sink`` (``LocoManipulationTask.action_sink()``): evaluation mode — the head epilogue also hands action = mean to the task."""
        n = obs.shape[0]
        if out is None:
            mu = torch.empty((n, 12), device=self._device, dtype=torch.float32)
            value = torch.empty((n, 1), device=self._device, dtype=torch.float32)
        else:
            mu, value = out
        stream = torch.cuda.current_stream(self._device).cuda_stream
        import ctypes as C
        with _lib.nvtx("mlp_forward"):
            rc = self._lib.lm_mlp_forward_sink(self._handle, obs.data_ptr(), _lib.ptr(norm_mean), _lib.ptr(norm_inv_std), float(norm_clip),
                                               mu.data_ptr(), value.data_ptr(), n, C.byref(sink) if sink is not None else None, stream)
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
