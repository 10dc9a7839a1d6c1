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

    def __init__(self, device="cuda:0", e_clip=0.2, critic_coef=2.0, entropy_coef=0.0, bounds_loss_coef=1e-3, clip_value=True,
                 stats_out: Optional[torch.Tensor] = None, grad_log_std_out: Optional[torch.Tensor] = None):
        """Defaults = cfg/train/JointLocomanipulationPPO.yaml:52-72.  ``stats_out`` [8] / ``grad_log_std_out`` [12]: let the kernel
        write straight into a caller-owned place (the tail of the gradient arena / the log_std gradient view)."""
        self.device = torch.device(device)
        self.cfg = dict(e_clip=float(e_clip), critic_coef=float(critic_coef), entropy_coef=float(entropy_coef),
                        bounds_loss_coef=float(bounds_loss_coef), clip_value=int(bool(clip_value)))
        self._lib, self._h = _handle(self.device)
        self.stats = stats_out if stats_out is not None else torch.zeros(8, device=self.device)
        self.grad_log_std = grad_log_std_out if grad_log_std_out is not None else torch.zeros(12, device=self.device)

    def __call__(self, mu, value, log_std, actions, old_neglogp, advantages, old_values, returns, old_mu=None, old_log_std=None,
                 out: Optional[Dict[str, torch.Tensor]] = None, update_old: bool = False):
        """All tensors fp32 contiguous on the device.  Returns (grad_mu [B,12], grad_value [B], grad_log_std [12], stats [8]).
        ``update_old``: rl_games' ``dataset.update_mu_sigma`` — after the KL has been taken, ``old_mu`` / ``old_log_std`` are
        overwritten in place with this pass's mu / log_std."""
        Bn = mu.shape[0]
        grad_mu = out["grad_mu"] if out else torch.empty(Bn, 12, device=self.device)
        grad_value = out["grad_value"] if out else torch.empty(Bn, device=self.device)
        p = _lib.ptr
        b = _lib.LmPpoBatch(p(mu), p(value), p(log_std), p(actions), p(old_neglogp), p(advantages), p(old_values), p(returns),
                            p(old_mu), p(old_log_std), Bn, self.cfg["e_clip"], self.cfg["critic_coef"], self.cfg["entropy_coef"],
                            self.cfg["bounds_loss_coef"], self.cfg["clip_value"], int(bool(update_old) and old_mu is not None))
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
    """torch.optim.Adam (+ clip_grad_norm_) over one flat fp32 arena; ``params`` / ``grads`` are 1-D views of it.  The learning
    rate lives on the device (``lr_dev``): with ``kl`` given to :meth:`step`, rl_games' adaptive schedule runs inside the launch
    (no host round trip per minibatch).  ``lr`` reads it back (host sync) / sets it."""

    def __init__(self, params: torch.Tensor, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 kl_threshold: Optional[float] = None, lr_bounds=(1e-6, 1e-2)):
        if not params.is_cuda:
            raise _lib.LmError("FlatAdam runs on CUDA tensors only (no CPU fallback)")
        self.params = params
        self.exp_avg = torch.zeros_like(params)
        self.exp_avg_sq = torch.zeros_like(params)
        self.lr_dev = torch.full((1,), float(lr), device=params.device)
        self.betas, self.eps, self.weight_decay, self.max_grad_norm = betas, float(eps), float(weight_decay), float(max_grad_norm)
        self.kl_threshold, self.lr_bounds = kl_threshold, lr_bounds
        self.step_count = 0
        self.grad_norm = torch.zeros(1, device=params.device)
        self._lib = _lib.load()

    @property
    def lr(self) -> float:
        return float(self.lr_dev.item())

    @lr.setter
    def lr(self, v: float) -> None:
        self.lr_dev.fill_(float(v))

    def step(self, grads: torch.Tensor, grad_scale: float = 1.0, kl: Optional[torch.Tensor] = None, kl_scale: float = 1.0) -> None:
        """``kl`` [1] (device): the minibatch KL, summed over ranks when ``kl_scale`` = 1 / world; moves the rate AFTER this step as
        ``AdaptiveScheduler.update`` does (> 2 thr: / 1.5, < thr / 2: x 1.5, clamped to ``lr_bounds``)."""
        self.step_count += 1
        stream = torch.cuda.current_stream(self.params.device).cuda_stream
        use_kl = kl is not None and self.kl_threshold is not None
        sched = _lib.LmAdamSchedule(self.lr_dev.data_ptr(), kl.data_ptr() if use_kl else None, float(kl_scale),
                                    float(self.kl_threshold or 0.0), float(self.lr_bounds[0]), float(self.lr_bounds[1]))
        rc = self._lib.lm_adam_step_scheduled(self.params.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                              self.params.numel(), C.byref(sched), 0.0, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                              self.step_count, self.max_grad_norm, float(grad_scale), self.grad_norm.data_ptr(), stream)
        _lib.check(rc, "lm_adam_step")


def _mlp_params(m):
    l1, l2, l3 = m.net[0], m.net[2], m.net[4]
    return [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, m.mean_layer.weight, m.mean_layer.bias, m.value_layer.weight, m.value_layer.bias]


class MlpPpoGradients:
    """Gradients of the rl_games PPO loss for ``MlpActorCritic`` on tcgen05 end to end (csrc/lm_mlp.cu, lm_ppo.cu, lm_mlp_train.cu):

        training forward (activations saved as bf16 tile blobs) -> lm_ppo_loss -> backward through the activations ->
        weight / bias gradients accumulated in TMEM over the minibatch [-> pushed to the peer ranks and summed: ``comm``]

    bf16 operands with fp32 accumulation throughout (what torch autocast would run); the result lands in the parameters' ``.grad``
    tensors, ready for ``FlatAdam``.  With ``arena`` (the flat buffer all ``.grad`` are views of, + an 8-float tail that receives the
    loss head's statistics) and ``comm`` (``dist.PeerComm``), the final reduction of the weight-gradient kernel is fused with the
    data-parallel all-reduce over NVLink peer memory (``lm_mlp_wgrad_allreduce``): on return the arena holds the SUM over ranks.
    """

    def __init__(self, model, max_batch: int, arena: Optional[torch.Tensor] = None, comm=None, **loss_cfg):
        self.model, self.device = model, model._device
        self.max_batch = int(max_batch)
        self.arena, self.comm = arena, comm
        self._lib = _lib.load()
        model._ensure()
        params = _mlp_params(model)
        for q in params + [model.log_std_parameter]:
            if q.grad is None:
                q.grad = torch.zeros_like(q)
        stats_out = None
        if arena is not None:
            n_par = sum(q.numel() for q in params) + 12
            if arena.numel() != n_par + 8:
                raise _lib.LmError("arena must hold every gradient plus an 8-float tail")
            stats_out = arena[n_par:]
            base, end = arena.data_ptr(), arena.data_ptr() + 4 * arena.numel()
            for q in params + [model.log_std_parameter]:
                if not (base <= q.grad.data_ptr() and q.grad.data_ptr() + 4 * q.numel() <= end):
                    raise _lib.LmError("every .grad must be a view of the arena")
            self._ls_off = (model.log_std_parameter.grad.data_ptr() - base) // 4
            self._tail_off = n_par
        self.loss = PpoLoss(device=self.device, stats_out=stats_out,
                            grad_log_std_out=model.log_std_parameter.grad if arena is not None else None, **loss_cfg)
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
                 norm_mean=None, norm_inv_std=None, norm_clip: float = 5.0, update_old: bool = False):
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
        _, _, g_ls, stats = self.loss(mu, value, log_std, actions, old_neglogp, advantages, old_values, returns, old_mu, old_log_std, out=out,
                                      update_old=update_old)
        _lib.check(lib.lm_mlp_backward(m._handle, p(out["grad_mu"]), p(out["grad_value"]), p(self.a1), p(self.a2), p(self.a3),
                                       p(self.dh), p(self.dz3), p(self.dz2), p(self.dz1), n, s), "lm_mlp_backward")
        grads = [q.grad.data_ptr() for q in _mlp_params(m)]
        blobs = (p(self.dz1), p(self.dz2), p(self.dz3), p(self.dh), p(self.a0), p(self.a1), p(self.a2), p(self.a3))
        if self.comm is not None and self.comm.world > 1:
            _lib.check(lib.lm_mlp_wgrad_allreduce(self.comm.handle, *blobs, n, p(self._ws), self._ctas, *grads, p(self.arena), self.arena.numel(),
                                                  self._ls_off, 12, self._tail_off, 8, s), "lm_mlp_wgrad_allreduce")
        else:
            _lib.check(lib.lm_mlp_wgrad(*blobs, n, p(self._ws), self._ctas, *grads, s), "lm_mlp_wgrad")
        if g_ls.data_ptr() != m.log_std_parameter.grad.data_ptr():
            m.log_std_parameter.grad.copy_(g_ls)
        return stats


class GnnPpoGradients:
    """The same for ``GnnActorCritic``: forward on the fused kernels (the module's ``precision``) -> ``lm_ppo_loss`` ->
    ``lm_gnn_backward`` (csrc/lm_gnn_train.cu: one kernel recomputes the fp32 forward per env, walks the limb graph backwards and
    accumulates all 20 weight / bias gradients per CTA in shared memory).  With ``comm`` the flat arena (gradients + the loss
    head's statistics) is summed over the ranks by the one-shot NVLink all-reduce right behind it."""

    def __init__(self, model, max_batch: int, arena: Optional[torch.Tensor] = None, comm=None, **loss_cfg):
        self.model, self.device = model, model._device
        self.max_batch, self.arena, self.comm = int(max_batch), arena, comm
        self._lib = _lib.load()
        model._ensure()
        params = model.ordered_parameters()
        for q in params + [model.log_std_parameter]:
            if q.grad is None:
                q.grad = torch.zeros_like(q)
        stats_out = None
        if arena is not None:
            n_par = sum(q.numel() for q in params) + 12
            if arena.numel() != n_par + 8:
                raise _lib.LmError("arena must hold every gradient plus an 8-float tail")
            stats_out = arena[n_par:]
        self.loss = PpoLoss(device=self.device, stats_out=stats_out,
                            grad_log_std_out=model.log_std_parameter.grad if arena is not None else None, **loss_cfg)
        self.mu = torch.empty(self.max_batch, 12, device=self.device)
        self.value = torch.empty(self.max_batch, 1, device=self.device)
        self.grad_mu = torch.empty(self.max_batch, 12, device=self.device)
        self.grad_value = torch.empty(self.max_batch, device=self.device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ctas = self._lib.lm_device_sm_count(idx)
        self._ws = torch.empty(self._ctas * self._lib.lm_gnn_grad_workspace_floats(), device=self.device)

    @torch.no_grad()
    def __call__(self, obs, actions, old_neglogp, advantages, old_values, returns, old_mu=None, old_log_std=None,
                 norm_mean=None, norm_inv_std=None, norm_clip: float = 5.0, update_old: bool = False):
        if norm_mean is not None:
            raise _lib.LmError("the GNN kernels take pre-processed observations (skrl applies its state scaler outside the model)")
        m, lib = self.model, self._lib
        n = obs.shape[0]
        if n > self.max_batch:
            raise _lib.LmError(f"minibatch of {n} exceeds max_batch {self.max_batch}")
        mu, value = self.mu[:n], self.value[:n]
        m(obs, out=(mu, value))
        s = torch.cuda.current_stream(self.device).cuda_stream
        out = {"grad_mu": self.grad_mu[:n], "grad_value": self.grad_value[:n]}
        _, _, g_ls, stats = self.loss(mu, value.view(-1), m.log_std_parameter.detach(), actions, old_neglogp, advantages, old_values, returns,
                                      old_mu, old_log_std, out=out, update_old=update_old)
        grads = m.ordered_parameters()
        arr = (C.c_void_p * 20)(*[q.grad.data_ptr() for q in grads])
        _lib.check(lib.lm_gnn_backward(m._handle, obs.data_ptr(), out["grad_mu"].data_ptr(), out["grad_value"].data_ptr(), n, self._ws.data_ptr(),
                                       self._ctas, arr, 20, s), "lm_gnn_backward")
        if g_ls.data_ptr() != m.log_std_parameter.grad.data_ptr():
            m.log_std_parameter.grad.copy_(g_ls)
        if self.comm is not None and self.comm.world > 1 and self.arena is not None:
            self.comm.all_reduce(self.arena)
        return stats


class PpoUpdater:
    """rl_games' ``train_epoch`` for the MLP actor-critic on the library's kernels (cfg/train/JointLocomanipulationPPO.yaml:44-72):

    * ``prepare_dataset``: GAE over the time-major rollout (``lm_gae``); ``normalize_value``: the value normaliser is updated with
      the values batch and then with the returns batch, each normalised right after its update (``lm_rms_update`` /
      ``lm_normalize_clamp``); ``normalize_advantage``: (a - mean) / (std + 1e-8) from ``lm_moments`` -> ``lm_moment_normalize``
      (``global_advantage_stats``: the three moments are summed over the ranks first — SURVEY 8e (3); rl_games 1.6.0 itself
      normalises per rank);
    * ``mini_epochs`` passes over contiguous minibatches (rl_games' PPODataset slices the flattened batch without shuffling, so
      every minibatch is a zero-copy view of the rollout buffers); per minibatch: ``normalize_input`` folds the minibatch's
      observations into the running statistics first (the train-mode forward of rl_games' model), tcgen05 forward / loss head /
      backward / weight gradients, the data-parallel gradient all-reduce — fused into the weight-gradient reduction over NVLink
      peer memory when ``comm`` is a ``dist.PeerComm``, one ``torch.distributed.all_reduce`` of the flat arena otherwise —, clip +
      Adam, re-pack of the bf16 operand blobs, ``dataset.update_mu_sigma`` (in the loss kernel);
    * ``lr_schedule: adaptive`` with rl_games' default ``schedule_type: legacy``: the rate moves after EVERY minibatch from that
      minibatch's KL averaged over the ranks.  The KL rides in the tail of the all-reduced gradient arena and the decision is
      taken inside the Adam launch, identically on every rank: no host synchronisation and no divergence of the replicas.
      ``schedule_type="standard"``: once per mini-epoch from the mean KL (one host sync per mini-epoch).
    """

    def __init__(self, model, rollout, minibatch_size: int, mini_epochs: int = 5, lr: float = 3e-4, gamma: float = 0.99, tau: float = 0.95,
                 e_clip: float = 0.2, critic_coef: float = 2.0, entropy_coef: float = 0.0, bounds_loss_coef: float = 1e-3, clip_value: bool = True,
                 max_grad_norm: float = 1.0, normalize_advantage: bool = True, kl_threshold: Optional[float] = 0.008, lr_bounds=(1e-6, 1e-2),
                 schedule_type: str = "legacy", normalize_value: bool = False, obs_rms=None, comm=None, global_advantage_stats: bool = False):
        from .rollout import RunningMeanStd
        self.model, self.rollout = model, rollout
        T, N = rollout.horizon, rollout.num_envs
        if (T * N) % minibatch_size:
            raise ValueError("batch size must be divisible by the minibatch size (rl_games asserts the same)")
        if schedule_type not in ("legacy", "standard"):
            raise ValueError(schedule_type)
        self.minibatch_size, self.mini_epochs = int(minibatch_size), int(mini_epochs)
        self.gamma, self.tau, self.normalize_advantage = float(gamma), float(tau), bool(normalize_advantage)
        self.kl_threshold, self.lr_bounds, self.schedule_type = kl_threshold, lr_bounds, schedule_type
        self.comm, self.global_advantage_stats = comm, bool(global_advantage_stats)
        dev = rollout.device
        self._lib = _lib.load()
        # parameters and gradients as views of two flat arenas (the layout FlatAdam and the fused all-reduce work on); the gradient
        # arena's 8-float tail receives the loss head's statistics, so the KL crosses the ranks inside the gradient message
        params = list(model.parameters())
        self.flat = torch.cat([p.detach().reshape(-1) for p in params])
        n_par = self.flat.numel()
        self.grad_storage = torch.zeros(n_par + 8, device=dev)
        self.flat_grad, self.tail = self.grad_storage[:n_par], self.grad_storage[n_par:]
        off = 0
        for p in params:
            k = p.numel()
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k
        model.mark_updated()
        self.opt = FlatAdam(self.flat, lr=lr, max_grad_norm=max_grad_norm, kl_threshold=kl_threshold, lr_bounds=lr_bounds)
        from .policy.gnn import GnnActorCritic
        engine = GnnPpoGradients if isinstance(model, GnnActorCritic) else MlpPpoGradients     # the GNN policy trains on its own kernels
        self.grads = engine(model, max_batch=self.minibatch_size, arena=self.grad_storage, comm=comm, e_clip=e_clip, critic_coef=critic_coef,
                            entropy_coef=entropy_coef, bounds_loss_coef=bounds_loss_coef, clip_value=clip_value)
        self.advantages = torch.empty(T, N, device=dev)
        self.returns = torch.empty(T, N, device=dev)
        self.old_values = torch.empty(T, N, device=dev)
        nmb = (T * N) // self.minibatch_size
        self.old_log_std = torch.empty(nmb, 12, device=dev)       # rl_games keeps sigma per sample; it is state-independent here: one row per minibatch
        self._moments = torch.zeros(3, dtype=torch.float64, device=dev)
        self._kl_acc = torch.zeros(1, device=dev)
        self.value_rms = RunningMeanStd(1, device=dev) if normalize_value else None
        self.obs_rms = obs_rms
        self.value_denorm = self.value_rms.export_denorm() if normalize_value else None     # what RolloutBuffer.write(value_denorm=...) takes
        from .rollout import _handle
        _, self._ws = _handle(dev)

    def _world(self) -> int:
        if self.comm is not None:
            return self.comm.world
        import torch.distributed as tdist
        return tdist.get_world_size() if (tdist.is_available() and tdist.is_initialized()) else 1

    @torch.no_grad()
    def update(self, last_values: torch.Tensor, last_dones: torch.Tensor, norm_mean=None, norm_inv_std=None):
        """One PPO update over the rollout currently in the buffer.  Returns the last minibatch's loss-head stats (device tensor [8];
        summed over ranks when sharded — divide by the world size)."""
        with _lib.nvtx("ppo_update"):
            return self._update(last_values, last_dones, norm_mean, norm_inv_std)

    def _update(self, last_values, last_dones, norm_mean, norm_inv_std):
        import torch.distributed as tdist
        from . import gae as _gae
        lib, s = self._lib, torch.cuda.current_stream(self.rollout.device).cuda_stream
        td = self.rollout.tensor_dict
        T, N = self.rollout.horizon, self.rollout.num_envs
        world = self._world()
        adv, ret = _gae.compute_gae(td["rewards"], td["values"], td["dones"], last_values, last_dones, gamma=self.gamma, lam=self.tau,
                                    variant=_gae.RLGAMES)
        if self.value_rms is not None:                    # prepare_dataset: values, then returns, through value_mean_std in train mode
            self.value_rms.update(td["values"].reshape(-1, 1))
            self.value_rms.normalize(td["values"], out=self.old_values)
            self.value_rms.update(ret.reshape(-1, 1))
            self.value_rms.normalize(ret, out=self.returns)
            self.value_rms.export_denorm()                # self.value_denorm (same storage) now holds the statistics the next rollout de-normalises with
        else:
            self.old_values.copy_(td["values"])
            self.returns.copy_(ret)
        if self.normalize_advantage:
            _lib.check(lib.lm_moments(self._ws, adv.data_ptr(), adv.numel(), self._moments.data_ptr(), s), "lm_moments")
            if self.global_advantage_stats and world > 1:
                tdist.all_reduce(self._moments)
            _lib.check(lib.lm_moment_normalize(adv.data_ptr(), self.advantages.data_ptr(), adv.numel(), self._moments.data_ptr(), 1e-8, s),
                       "lm_moment_normalize")
        else:
            self.advantages.copy_(adv)
        self.old_log_std.copy_(self.model.log_std_parameter.detach().expand_as(self.old_log_std))
        flat = lambda t: t.view(T * N, *t.shape[2:])
        obs, actions, mus = flat(td["obses"]), flat(td["actions"]), flat(td["mus"])
        values, neglogp = flat(self.old_values), flat(td["neglogpacs"])
        advs, rets = flat(self.advantages), flat(self.returns)
        fused = self.comm is not None and self.comm.world > 1
        legacy = self.schedule_type == "legacy" and self.kl_threshold is not None
        stats = None
        nmb = (T * N) // self.minibatch_size
        for _ in range(self.mini_epochs):
            self._kl_acc.zero_()
            for i in range(nmb):
                sl = slice(i * self.minibatch_size, (i + 1) * self.minibatch_size)
                if self.obs_rms is not None:              # rl_games' model normalises inside a train-mode forward: update, then use
                    self.obs_rms.update(obs[sl])
                    norm_mean, norm_inv_std = self.obs_rms.export()
                stats = self.grads(obs[sl], actions[sl], neglogp[sl], advs[sl], values[sl], rets[sl], mus[sl], self.old_log_std[i],
                                   norm_mean=norm_mean, norm_inv_std=norm_inv_std, update_old=True)
                if world > 1 and not fused:               # one collective per minibatch over the flat gradient arena + statistics tail
                    tdist.all_reduce(self.grad_storage)
                self.opt.step(self.flat_grad, grad_scale=1.0 / world, kl=self.tail[4:5] if legacy else None, kl_scale=1.0 / world)
                self.model.pack()                         # the kernels read the bf16 blobs, the optimiser wrote the fp32 masters
                if not legacy:
                    self._kl_acc += self.tail[4]
            if self.kl_threshold is not None and not legacy:        # 'standard': once per mini-epoch (one host sync)
                kl = float(self._kl_acc) / nmb / world
                lo, hi = self.lr_bounds
                lr = self.opt.lr
                if kl > 2.0 * self.kl_threshold:
                    lr = max(lr / 1.5, lo)
                if kl < 0.5 * self.kl_threshold:
                    lr = min(lr * 1.5, hi)
                self.opt.lr = lr
        return stats
