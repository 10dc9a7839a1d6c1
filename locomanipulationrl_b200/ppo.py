"""PPO update pieces on the GPU (csrc/lm_ppo.cu): the rl_games loss head in one launch and a fused clip + Adam step.

``PpoLoss`` returns the gradients of the rl_games ``calc_gradients`` loss with respect to the network outputs; the caller
back-propagates them through its own network (``torch.autograd.backward([mu, value], [grad_mu, grad_value])``).
``FlatAdam`` steps a flat fp32 parameter arena (the layout ``dist.FusedHorizonReducer`` keeps for the gradient all-reduce).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .rollout import _handle


class PpoLoss:
    STAT_NAMES = ("a_loss", "c_loss", "b_loss", "entropy", "kl", "loss", "clip_frac")

    def __init__(self, device="cuda:0", e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-4, clip_value=True):
        self.device = torch.device(device)
        self.cfg = dict(e_clip=float(e_clip), critic_coef=float(critic_coef), entropy_coef=float(entropy_coef),
                        bounds_loss_coef=float(bounds_loss_coef), clip_value=int(bool(clip_value)))
        self._lib, self._h = _handle(self.device)
        self.stats = torch.zeros(8, device=self.device)
        self.grad_log_std = torch.zeros(12, device=self.device)

    def __call__(self, mu, value, log_std, actions, old_neglogp, advantages, old_values, returns, old_mu=None, old_log_std=None,
                 out: Optional[Dict[str, torch.Tensor]] = None):
        """All tensors fp32 contiguous on the device.  Returns (grad_mu [B,12], grad_value [B], grad_log_std [12], stats [8])."""
        Bn = mu.shape[0]
        grad_mu = out["grad_mu"] if out else torch.empty(Bn, 12, device=self.device)
        grad_value = out["grad_value"] if out else torch.empty(Bn, device=self.device)
        p = _lib.ptr
        b = _lib.LmPpoBatch(p(mu), p(value), p(log_std), p(actions), p(old_neglogp), p(advantages), p(old_values), p(returns),
                            p(old_mu), p(old_log_std), Bn, self.cfg["e_clip"], self.cfg["critic_coef"], self.cfg["entropy_coef"],
                            self.cfg["bounds_loss_coef"], self.cfg["clip_value"], 0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.lm_ppo_loss(self._h, C.byref(b), grad_mu.data_ptr(), grad_value.data_ptr(), self.grad_log_std.data_ptr(),
                                   self.stats.data_ptr(), stream)
        _lib.check(rc, "lm_ppo_loss")
        return grad_mu, grad_value, self.grad_log_std, self.stats

    def __del__(self):
        try:
            if self._h:
                self._lib.lm_rollout_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FlatAdam:
    """torch.optim.Adam (+ clip_grad_norm_) over one flat fp32 arena; ``params`` / ``grads`` are 1-D views of it."""

    def __init__(self, params: torch.Tensor, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0):
        if not params.is_cuda:
            raise _lib.LmError("FlatAdam runs on CUDA tensors only (no CPU fallback)")
        self.params = params
        self.exp_avg = torch.zeros_like(params)
        self.exp_avg_sq = torch.zeros_like(params)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = float(lr), betas, float(eps), float(weight_decay), float(max_grad_norm)
        self.step_count = 0
        self.grad_norm = torch.zeros(1, device=params.device)
        self._lib = _lib.load()

    def step(self, grads: torch.Tensor, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        stream = torch.cuda.current_stream(self.params.device).cuda_stream
        rc = self._lib.lm_adam_step(self.params.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                    self.params.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    self.step_count, self.max_grad_norm, float(grad_scale), self.grad_norm.data_ptr(), stream)
        _lib.check(rc, "lm_adam_step")


class MlpPpoGradients:
    """Gradients of the rl_games PPO loss for ``MlpActorCritic`` on tcgen05 end to end (csrc/lm_mlp.cu, lm_ppo.cu, lm_mlp_train.cu):

        training forward (activations saved as bf16 tile blobs) -> lm_ppo_loss -> backward through the activations ->
        weight / bias gradients accumulated in TMEM over the minibatch

    bf16 operands with fp32 accumulation throughout (what torch autocast would run); the result lands in the parameters' ``.grad``
    tensors, ready for the fused all-reduce and ``FlatAdam``.
    """

    def __init__(self, model, max_batch: int, **loss_cfg):
        self.model, self.device = model, model._device
        self.max_batch = int(max_batch)
        self.loss = PpoLoss(device=self.device, **loss_cfg)
        self._lib = _lib.load()
        model._ensure()
        nb = lambda d: torch.empty(self._lib.lm_mlp_blob_bytes(self.max_batch, d), dtype=torch.uint8, device=self.device)
        self.a0, self.a1, self.a2, self.a3 = nb(64), nb(256), nb(128), nb(64)
        self.dz1, self.dz2, self.dz3, self.dh = nb(256), nb(128), nb(64), nb(16)
        self.mu = torch.empty(self.max_batch, 12, device=self.device)
        self.value = torch.empty(self.max_batch, device=self.device)
        self.grad_mu = torch.empty(self.max_batch, 12, device=self.device)
        self.grad_value = torch.empty(self.max_batch, device=self.device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ctas = self._lib.lm_device_sm_count(idx)
        self._ws = torch.empty(self._ctas * self._lib.lm_mlp_wgrad_workspace_floats(), device=self.device)

    @torch.no_grad()
    def __call__(self, obs, actions, old_neglogp, advantages, old_values, returns, old_mu=None, old_log_std=None,
                 norm_mean=None, norm_inv_std=None, norm_clip: float = 5.0):
        """One minibatch (all tensors fp32, contiguous, on the device).  Fills ``.grad`` of every parameter; returns the loss head's stats [8]."""
        m, lib = self.model, self._lib
        n = obs.shape[0]
        if n > self.max_batch:
            raise _lib.LmError(f"minibatch of {n} exceeds max_batch {self.max_batch}")
        if m._packed_version != m._param_version():
            m.pack()
        s = torch.cuda.current_stream(self.device).cuda_stream
        p = _lib.ptr
        mu, value = self.mu[:n], self.value[:n]
        _lib.check(lib.lm_mlp_forward_train(m._handle, p(obs), p(norm_mean), p(norm_inv_std), float(norm_clip), p(mu), p(value),
                                            p(self.a0), p(self.a1), p(self.a2), p(self.a3), n, s), "lm_mlp_forward_train")
        log_std = m.log_std_parameter.detach()
        out = {"grad_mu": self.grad_mu[:n], "grad_value": self.grad_value[:n]}
        _, _, g_ls, stats = self.loss(mu, value, log_std, actions, old_neglogp, advantages, old_values, returns, old_mu, old_log_std, out=out)
        _lib.check(lib.lm_mlp_backward(m._handle, p(out["grad_mu"]), p(out["grad_value"]), p(self.a1), p(self.a2), p(self.a3),
                                       p(self.dh), p(self.dz3), p(self.dz2), p(self.dz1), n, s), "lm_mlp_backward")
        l1, l2, l3 = m.net[0], m.net[2], m.net[4]
        params = [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, m.mean_layer.weight, m.mean_layer.bias, m.value_layer.weight, m.value_layer.bias]
        for q in params + [m.log_std_parameter]:
            if q.grad is None:
                q.grad = torch.zeros_like(q)
        _lib.check(lib.lm_mlp_wgrad(p(self.dz1), p(self.dz2), p(self.dz3), p(self.dh), p(self.a0), p(self.a1), p(self.a2), p(self.a3), n,
                                    p(self._ws), self._ctas, *[q.grad.data_ptr() for q in params], s), "lm_mlp_wgrad")
        m.log_std_parameter.grad.copy_(g_ls)
        return stats


class PpoUpdater:
    """rl_games' ``train_epoch`` for the MLP actor-critic on the library's kernels: GAE over the time-major rollout, advantage
    normalisation, ``mini_epochs`` passes over contiguous minibatches (rl_games' PPODataset slices the flattened batch without
    shuffling, so every minibatch is a zero-copy view of the rollout buffers), tcgen05 gradients, one all-reduce of the flat gradient arena when a
    process group exists, clip + Adam, and the KL-adaptive learning rate (cfg/train/JointLocomanipulationPPO.yaml:52-72).

    Not covered: rl_games' ``normalize_value`` (values / returns standardised by a running mean/std, cfg/train/*PPO.yaml:47) — the
    statistics kernel exists (``rollout.RunningMeanStd`` with dim 1) but is not wired into this loop; observation normalisation is
    passed in as ``norm_mean`` / ``norm_inv_std``.

    Tested single-process against the all-torch loop (tests/test_mlp_train.py); the ``world > 1`` all-reduce branch is checked by
    tools/ppo_updater_2rank.py under torchrun on 2 GPUs (identical parameters on both ranks, different from a rank on its own).
    """

    def __init__(self, model, rollout, minibatch_size: int, mini_epochs: int = 5, lr: float = 3e-4, gamma: float = 0.99, tau: float = 0.95,
                 e_clip: float = 0.2, critic_coef: float = 2.0, entropy_coef: float = 0.0, bounds_loss_coef: float = 1e-4, clip_value: bool = True,
                 max_grad_norm: float = 1.0, normalize_advantage: bool = True, kl_threshold: Optional[float] = 0.008, lr_bounds=(1e-6, 1e-2)):
        self.model, self.rollout = model, rollout
        T, N = rollout.horizon, rollout.num_envs
        if (T * N) % minibatch_size:
            raise ValueError("batch size must be divisible by the minibatch size (rl_games asserts the same)")
        self.minibatch_size, self.mini_epochs = int(minibatch_size), int(mini_epochs)
        self.gamma, self.tau, self.normalize_advantage = float(gamma), float(tau), bool(normalize_advantage)
        self.kl_threshold, self.lr_bounds = kl_threshold, lr_bounds
        dev = rollout.device
        # parameters and gradients as views of two flat arenas (the layout FlatAdam and the fused all-reduce work on)
        params = list(model.parameters())
        self.flat = torch.cat([p.detach().reshape(-1) for p in params])
        self.flat_grad = torch.zeros_like(self.flat)
        off = 0
        for p in params:
            k = p.numel()
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k
        self.opt = FlatAdam(self.flat, lr=lr, max_grad_norm=max_grad_norm)
        self.grads = MlpPpoGradients(model, max_batch=self.minibatch_size, e_clip=e_clip, critic_coef=critic_coef, entropy_coef=entropy_coef,
                                     bounds_loss_coef=bounds_loss_coef, clip_value=clip_value)
        self.advantages = torch.empty(T, N, device=dev)
        self.returns = torch.empty(T, N, device=dev)
        self.old_log_std = torch.empty(12, device=dev)
        self._kl = torch.zeros(1, device=dev)

    @torch.no_grad()
    def update(self, last_values: torch.Tensor, last_dones: torch.Tensor, norm_mean=None, norm_inv_std=None):
        """One PPO update over the rollout currently in the buffer.  Returns the last minibatch's loss-head stats (device tensor [8])."""
        from . import gae as _gae
        td = self.rollout.tensor_dict
        T, N = self.rollout.horizon, self.rollout.num_envs
        adv, ret = _gae.compute_gae(td["rewards"], td["values"], td["dones"], last_values, last_dones, gamma=self.gamma, lam=self.tau,
                                    variant=_gae.RLGAMES)
        self.returns.copy_(ret)
        if self.normalize_advantage:                      # rl_games: (adv - mean) / (std + 1e-8) over the whole batch
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        self.advantages.copy_(adv)
        self.old_log_std.copy_(self.model.log_std_parameter.detach())
        flat = lambda t: t.view(T * N, *t.shape[2:])
        obs, actions, mus = flat(td["obses"]), flat(td["actions"]), flat(td["mus"])
        values, neglogp = flat(td["values"]), flat(td["neglogpacs"])
        advs, rets = flat(self.advantages), flat(self.returns)
        import torch.distributed as tdist
        world = tdist.get_world_size() if (tdist.is_available() and tdist.is_initialized()) else 1
        stats = None
        for _ in range(self.mini_epochs):
            self._kl.zero_()
            nmb = (T * N) // self.minibatch_size
            for i in range(nmb):
                sl = slice(i * self.minibatch_size, (i + 1) * self.minibatch_size)
                stats = self.grads(obs[sl], actions[sl], neglogp[sl], advs[sl], values[sl], rets[sl], mus[sl], self.old_log_std,
                                   norm_mean=norm_mean, norm_inv_std=norm_inv_std)
                if world > 1:                             # one collective per minibatch over the flat gradient arena (sum; scaled in the Adam step)
                    tdist.all_reduce(self.flat_grad)
                self.opt.step(self.flat_grad, grad_scale=1.0 / world)
                self._kl += stats[4]
            if self.kl_threshold is not None:             # rl_games adaptive schedule, once per mini-epoch (one host sync)
                kl = float(self._kl) / nmb
                lo, hi = self.lr_bounds
                if kl > 2.0 * self.kl_threshold:
                    self.opt.lr = max(self.opt.lr / 1.5, lo)
                elif kl < 0.5 * self.kl_threshold:
                    self.opt.lr = min(self.opt.lr * 1.5, hi)
        return stats
