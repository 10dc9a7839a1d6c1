"""Reference-signature task classes: ``Task(name, sim_config, env, offset=None)`` as built by
``initialize_task`` (utils/task_util.py:60-94), reading the same config keys as the reference constructors
(QPC:108-115; rl_task.py:67-79)."""
from __future__ import annotations

from typing import Optional

from ..configs import TASK_CONFIGS, get_task_config
from .base import LocoManipulationTask


def config_from_sim_config(name: str, sim_config):
    """Overlay the yaml values the reference reads onto the task's class constants."""
    cfg = get_task_config(name)
    task_cfg, root = sim_config.task_config, sim_config.config
    env = root.get("task", {}).get("env", {})
    return cfg.with_(dt=task_cfg["sim"]["dt"], max_episode_length=task_cfg["sim"]["max_episode_length"],
                     env_spacing=task_cfg["env"].get("envSpacing", cfg.env_spacing),
                     clip_observations=env.get("clipObservations", cfg.clip_observations),
                     clip_actions=env.get("clipActions", cfg.clip_actions),
                     control_frequency_inv=env.get("controlFrequencyInv", cfg.control_frequency_inv)), task_cfg["env"]["numEnvs"]


def make_reference_task_class(task_name: str):
    if task_name not in TASK_CONFIGS:
        raise KeyError(task_name)

    class _Task(LocoManipulationTask):
        def __init__(self, name, sim_config, env, offset=None, scene=None):
            cfg, num_envs = config_from_sim_config(task_name, sim_config)
            self._sim_config, self._cfg, self._task_cfg, self._env = sim_config, sim_config.config, sim_config.task_config, env
            device = sim_config.config.get("sim_device", "cuda:0")
            super().__init__(cfg, num_envs, scene=scene, device=device, rl_device=sim_config.config.get("rl_device", device))
            self.test = sim_config.config.get("test", False)
            dr = sim_config.task_config.get("domain_randomization") or {}
            params = dr.get("randomization_params") or {}
            if dr.get("randomize", False) and (params.get("observations") or params.get("actions")):     # randomize.py:41-56,99-104
                self.enable_noise_randomization(params, seed=int(sim_config.config.get("seed", 0)))

    _Task.__name__ = task_name
    _Task.__qualname__ = task_name
    return _Task


TASK_MAP = {name: make_reference_task_class(name) for name in TASK_CONFIGS}
