"""rl_games / skrl hooks for the two policies.

The reference registers only the vec-env with rl_games (scripts/rlgames_train.py:58-63) and wires the GNN through
skrl ``Model`` mixins (scripts/skrl_ppo_locomanipulation.py:25-61).  Neither library is installed in this image, so the
adapters import them lazily; the ``forward`` / ``compute`` bodies themselves only need torch + liblm_b200.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .gnn import GnnActorCritic
from .mlp import MlpActorCritic


def _core_forward(core, obs):
    """Inference: the fused kernels.  With autograd on (rl_games ``calc_gradients`` / skrl ``_update``): the MLP's differentiable
    tcgen05 path (``_MlpTrainFunction``); the GNN's differentiable torch restatement (``reference_forward``) — its update kernels
    are ``ppo.GnnPpoGradients``, which works outside autograd."""
    if torch.is_grad_enabled() and isinstance(core, GnnActorCritic) and any(p.requires_grad for p in core.parameters()):
        return core.reference_forward(obs)
    return core(obs)


class _RlGamesNet(nn.Module):
    """rl_games 1.6.0 network contract: ``forward(obs_dict) -> (mu, sigma, value, states)``; ``is_rnn() -> False``."""

    def __init__(self, core: nn.Module, fixed_sigma: bool = True):
        super().__init__()
        self.core = core
        self.fixed_sigma = fixed_sigma

    def is_rnn(self):
        return False

    def is_separate_critic(self):
        return False

    def forward(self, obs_dict):
        obs = obs_dict["obs"]
        mu, value, log_std = _core_forward(self.core, obs)
        sigma = log_std.expand_as(mu) if self.fixed_sigma else log_std     # rl_games applies exp() in the model wrapper
        return mu, sigma, value, obs_dict.get("rnn_states", None)


class _Builder:
    """``NetworkBuilder`` API: ``load(params)`` then ``build(name, **kwargs)``."""
    core_cls = MlpActorCritic

    def __init__(self, **kwargs):
        self.params = {}

    def load(self, params):
        self.params = params or {}

    def build(self, name, **kwargs):
        actions_num = kwargs.get("actions_num", 12)
        input_shape = kwargs.get("input_shape", (64,))
        assert actions_num == 12 and tuple(input_shape) == (64,), "kernels are specialised to the tasks' 64 -> 12 network"
        device = self.params.get("device", "cuda:0")
        if self.core_cls is MlpActorCritic:
            core = MlpActorCritic(device=device)
        else:
            core = GnnActorCritic(device=device)
        return _RlGamesNet(core, fixed_sigma=self.params.get("space", {}).get("continuous", {}).get("fixed_sigma", True))

    def __call__(self, name, **kwargs):
        return self.build(name, **kwargs)


class RlGamesMlpBuilder(_Builder):
    core_cls = MlpActorCritic


class RlGamesGnnBuilder(_Builder):
    core_cls = GnnActorCritic


def register_rlgames_networks():
    """``model_builder.register_network('locomani_mlp' | 'locomani_gnn', Builder)`` (needs rl_games installed)."""
    from rl_games.algos_torch import model_builder  # noqa: WPS433 (optional dependency)
    model_builder.register_network("locomani_mlp", RlGamesMlpBuilder)
    model_builder.register_network("locomani_gnn", RlGamesGnnBuilder)


class SkrlShared(nn.Module):
    """Drop-in for the reference's ``Shared`` model body (scripts/skrl_ppo_locomanipulation.py:25-61): same ``compute``
    contract.  With skrl installed, mix it as ``class Shared(GaussianMixin, DeterministicMixin, Model, SkrlShared)``."""

    def __init__(self, use_graph: bool = False, device="cuda:0"):
        super().__init__()
        self.core = GnnActorCritic(device=device) if use_graph else MlpActorCritic(device=device)
        self._cache = None
        self._register_load_state_dict_pre_hook(self._remap_reference_keys)

    @staticmethod
    def _remap_reference_keys(state_dict, prefix, *args):
        """A checkpoint of the reference's ``Shared`` has its parameters at the top level (``net.0.weight``, ``mean_layer.weight``,
        ``log_std_parameter`` ...); here they live under ``core.``: accept both."""
        for k in list(state_dict.keys()):
            if k.startswith(prefix) and not k.startswith(prefix + "core."):
                state_dict[prefix + "core." + k[len(prefix):]] = state_dict.pop(k)

    def compute(self, inputs, role):
        states = inputs["states"]
        # The reference evaluates the shared trunk once per role; here both heads come out of one kernel launch, reused by the second
        # role of the SAME step.  The task hands out the same obs_buf object every step and the kernels rewrite it in place, so the
        # key is (storage, torch version counter, role already served) — never object identity alone — and differentiable calls
        # are never cached (each role's loss needs its own graph).
        key = (states.data_ptr(), states._version, tuple(states.shape))
        if torch.is_grad_enabled() or self._cache is None or self._cache[0] != key or role in self._cache[2]:
            self._cache = (key, _core_forward(self.core, states), set())
        self._cache[2].add(role)
        mu, value, log_std = self._cache[1]
        if role == "policy":
            return mu, log_std, {}
        if role == "value":
            return value, {}
        raise ValueError(role)
