"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the UNMODIFIED reference under the shim.

Run in the authoring container only (needs /root/reference):  ``python -m oracle.make_golden``.
Each task fixture stores, for a short seeded rollout, everything the reference task exposes after each
``pre_physics_step`` / ``post_physics_step`` so that (a) the port and (b) the CUDA path can be checked on a
box where the reference does not exist.  Inputs are regenerated from seeds (SyntheticScene, torch.Generator).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from locomanipulationrl_b200 import configs, synthetic
from oracle import refshim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TASK_N = 50            # ragged: n % 4 != 0 (co-train halves of 25) -> exercises the non-TMA tail path of the kernels
TASK_STEPS = 20        # long enough for the consecutive-success counter to pass 15 and trigger goal resets
FULL_STEPS = (0, 1, 2, 17, 18, 19)   # steps whose float tensors are stored; the rest keep masks / ints / rewards only
STEPS_OVERRIDE = {"QuadrupedPoseControlCustomController": 28, "QuadrupedManipulatePlateCustomController": 28}   # goal reset needs consecutive_successes > 20 (mirrored in tests/common.py)
SEED = 7


def rollout_inputs(n, steps, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    acts = [(torch.rand(n, 12, generator=g) * 2 - 1) for _ in range(steps)]
    unis = [torch.rand(n, 3, generator=g) for _ in range(steps)]
    return acts, unis


def task_fixture(name):
    cfg = configs.get_task_config(name)
    scene = synthetic.SyntheticScene(cfg, TASK_N, seed=SEED)
    task = refshim.make_reference_task(cfg, scene)
    steps = STEPS_OVERRIDE.get(name, TASK_STEPS)
    full_steps = (0, 1, 2, steps - 3, steps - 2, steps - 1)
    acts, unis = rollout_inputs(TASK_N, steps)
    out = {}
    pd = cfg.control_mode == configs.CONTROL_PD
    for t in range(steps):
        if pd:   # the PD loop steps the simulation itself: feed it the seeded sub-step joint states (generated before torch.rand is patched)
            sub = [scene.substep_joint_state(t, i) for i in range(cfg.pd_control_decimal)]
            cursor = {"i": 0}

            def hook():
                scene.load(sub[cursor["i"]])
                cursor["i"] += 1
            refshim._SimulationContext.hook = hook
        with refshim.supplied_uniforms(unis[t]), refshim.quiet():
            ids = task.reset_buf.nonzero(as_tuple=False).squeeze(-1).clone()
            task.pre_physics_step(acts[t].clone())
        refshim._SimulationContext.hook = None
        full = t in full_steps
        if pd:
            out[f"s{t}_torque"] = task.torque.numpy().copy()
            if full:
                out[f"s{t}_pd_targets"] = task.current_joint_position_targets.numpy().copy()
                out[f"s{t}_pd_targets_se"] = task.current_joint_position_targets_se.numpy().copy()
        out[f"s{t}_reset_ids"] = ids.numpy()
        out[f"s{t}_goal"] = task.goal_quaternions.numpy().copy()
        for grp in ("loco", "mani"):
            st = getattr(scene, grp)
            if st and full:
                art = (task.robot_locomotion if grp == "loco" else task.robot_manipulation)._robot_articulation
                out[f"s{t}_{grp}_targets"] = art.targets["effort" if pd else "velocity"][:, :12].numpy().copy()
                for k in ("dof_pos", "dof_vel", "indicator_pos", "indicator_quat") + (("root_pos", "root_quat") if grp == "loco" else ("plate_pos", "plate_quat")):
                    out[f"s{t}_{grp}_{k}"] = st[k].numpy().copy()
        scene.step(t, task.goal_quaternions)
        with refshim.quiet():
            obs, rew, reset, extras = task.post_physics_step()
        if full:
            out[f"s{t}_obs"] = obs.numpy().copy()
            if cfg.num_states == 93:
                out[f"s{t}_states"] = task.get_states().numpy().copy()
            out[f"s{t}_last_base_tip_positions"] = task.last_base_tip_positions.numpy().copy()
        out[f"s{t}_rew"] = rew.numpy().copy()
        out[f"s{t}_reset"] = reset.numpy().astype(np.int8)
        for k in ("progress_buf", "successes", "consecutive_successes", "goal_reset_buf"):
            out[f"s{t}_{k}"] = getattr(task, k).numpy().astype(np.int16)
        for k, v in extras.items():
            out[f"s{t}_extras/{k}"] = np.asarray(float(v), dtype=np.float32)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"task_{name}.npz"), **out)
    print("wrote", name, len(out), "arrays")


def gnn_fixture():
    refshim.install()
    import graph_model_orebot_ov as G  # the reference module (scripts/graph_model_orebot_ov.py)
    torch.manual_seed(11)
    net = G.GraphNet(hidden_features=32, out_features=32)
    act = G.Action_Layer(hidden_features=32, num_actions=12)
    val = G.Value_Layer(hidden_features=32)
    obs = torch.randn(64, 64) * 1.5
    with torch.no_grad():
        h = net(obs)
        mu, value = act(h), val(h)
    sd = {"net." + k: v for k, v in net.state_dict().items()}
    sd.update({"mean_layer." + k: v for k, v in act.state_dict().items()})
    sd.update({"value_layer." + k: v for k, v in val.state_dict().items()})
    out = {"obs": obs.numpy(), "mu": mu.numpy(), "value": value.numpy()}
    out.update({"p/" + k: v.numpy() for k, v in sd.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gnn_reference.npz"), **out)
    print("wrote gnn", mu.shape, value.shape)


def math_fixture():
    """utils/math.py helpers on seeded inputs (the SciPy float64 path)."""
    refshim.install()
    import utils.math as M
    g = torch.Generator().manual_seed(5)
    n = 512
    q = torch.randn(n, 4, generator=g); q = q / q.norm(dim=-1, keepdim=True) + torch.randn(n, 4, generator=g) * 1e-4
    q2 = torch.randn(n, 4, generator=g); q2 = q2 / q2.norm(dim=-1, keepdim=True)
    t = torch.randn(n, 3, generator=g)
    v = torch.randn(n, 8, 3, generator=g)
    out = {"q": q.numpy(), "q2": q2.numpy(), "t": t.numpy(), "v": v.numpy(),
           "transform": M.transform_vectors(q, t, v).numpy(), "inverse_transform": M.inverse_transform_vectors(q, t, v).numpy(),
           "rotate": M.rotate_orientations(q, q2).numpy(), "inverse_rotate": M.inverse_rotate_orientations(q, q2).numpy(),
           "main_example": M.inverse_rotate_orientations(torch.tensor([-0.5, -0.5, 0.5, 0.5]).repeat(2, 1),
                                                         torch.tensor([0.7071, 0, 0, 0.7071]).repeat(2, 1)).numpy()}   # math.py:214-218
    np.savez_compressed(os.path.join(GOLDEN_DIR, "math_reference.npz"), **out)
    print("wrote math")


NOISE_CASES = {
    # name: (buffer type, dim, randomization params in the reference yaml's format, cfg/task/JointLocomanipulation.yaml:171-191)
    "obs_gauss_add": ("observations", 64, {"on_reset": {"operation": "additive", "distribution": "gaussian", "distribution_parameters": [0.0, 0.0001]},
                                          "on_interval": {"frequency_interval": 1, "operation": "additive", "distribution": "gaussian",
                                                          "distribution_parameters": [0.0, 0.002]}}),
    "act_gauss_add": ("actions", 12, {"on_reset": {"operation": "additive", "distribution": "gaussian", "distribution_parameters": [0.0, 0.015]},
                                      "on_interval": {"frequency_interval": 1, "operation": "additive", "distribution": "gaussian",
                                                      "distribution_parameters": [0.0, 0.05]}}),
    "obs_uniform_scale_every3": ("observations", 64, {"on_reset": {"operation": "scaling", "distribution": "uniform", "distribution_parameters": [0.9, 1.1]},
                                                     "on_interval": {"frequency_interval": 3, "operation": "additive", "distribution": "uniform",
                                                                     "distribution_parameters": [-0.05, 0.05]}}),
    "act_loguniform_interval_only": ("actions", 12, {"on_interval": {"frequency_interval": 2, "operation": "scaling", "distribution": "loguniform",
                                                                    "distribution_parameters": [0.8, 1.25]}}),
}
NOISE_N, NOISE_STEPS = 37, 8


def noise_inputs(case, n=NOISE_N, steps=NOISE_STEPS, seed=SEED):
    _, dim, _ = NOISE_CASES[case]
    g = torch.Generator().manual_seed(seed + len(case))
    bufs = [torch.randn(n, dim, generator=g) for _ in range(steps)]
    resets = [(torch.rand(n, generator=g) < (1.0 if t == 0 else 0.15)).long() for t in range(steps)]     # step 0: reset storm (reset_buf starts as ones)
    zs = [(torch.randn(n, dim, generator=g), torch.randn(n, dim, generator=g)) for _ in range(steps)]    # standard-normal rows (reset, interval)
    us = [(torch.rand(n, dim, generator=g), torch.rand(n, dim, generator=g)) for _ in range(steps)]      # U[0,1) rows
    return bufs, resets, zs, us


def noise_fixture():
    """utils/domain_randomization/randomize.py:212-305 run by the reference Randomizer, its draws supplied from outside."""
    refshim.install()
    import importlib
    R = importlib.import_module("omniisaacgymenvs.utils.domain_randomization.randomize")

    class _Cfg:
        pass

    class _Task:
        num_observations, num_actions = 64, 12
        randomize_observations = randomize_actions = False

    out = {}
    for case, (kind, dim, params) in NOISE_CASES.items():
        sc = _Cfg()
        sc.task_config = {"env": {"numEnvs": NOISE_N}, "domain_randomization": {"randomize": True, "randomization_params": {kind: params}}}
        sc.config = {"rl_device": "cpu", "seed": 0}
        rz = R.Randomizer(sc)
        task = _Task()
        (rz._set_up_observations_randomization if kind == "observations" else rz._set_up_actions_randomization)(task)
        bufs, resets, zs, us = noise_inputs(case)
        real_normal, real_rand = torch.normal, torch.rand
        for t in range(NOISE_STEPS):
            cursor = {"n": 0, "u": 0}       # first draw of a step = on_reset rows, second = on_interval rows (randomize.py order)

            def fake_normal(mean, std, size, device=None, _t=t, _c=cursor):
                z = zs[_t][_c["n"]][:size[0]]
                _c["n"] += 1
                return z * std + mean

            def fake_rand(size, device=None, _t=t, _c=cursor):
                u = us[_t][_c["u"]][:size[0]]
                _c["u"] += 1
                return u.clone()
            has_reset = "on_reset" in params
            torch.normal, torch.rand = fake_normal, fake_rand
            try:
                if not has_reset:
                    cursor["n"] = cursor["u"] = 1                    # the interval draw is the first call then: keep it on the interval rows
                buf = bufs[t].clone()
                if has_reset and int(resets[t].sum()) == 0:
                    cursor["n"] = cursor["u"] = 1                    # no reset draw this step (randomize.py:285 guards on len(reset_ids))
                fn = rz.apply_observations_randomization if kind == "observations" else rz.apply_actions_randomization
                res = fn(buf, resets[t])
            finally:
                torch.normal, torch.rand = real_normal, real_rand
            out[f"{case}/s{t}_out"] = res.numpy().copy()
            cnt = rz._observations_counter_buffer if kind == "observations" else rz._actions_counter_buffer
            cor = rz._observations_correlated_noise if kind == "observations" else rz._actions_correlated_noise
            out[f"{case}/s{t}_counter"] = cnt.numpy().astype(np.int16)
            out[f"{case}/s{t}_correlated"] = cor.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "noise_reference.npz"), **out)
    print("wrote noise", len(out), "arrays")


def main():
    if "--noise-only" in sys.argv:
        noise_fixture()
        return
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--task=")]
    if only:                                            # regenerate selected task fixtures only
        for name in only:
            task_fixture(name)
        return
    if not refshim.reference_available():
        sys.exit("reference tree not available; golden vectors can only be generated in the authoring container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in configs.TASK_CONFIGS:
        task_fixture(name)
    gnn_fixture()
    math_fixture()
    noise_fixture()


if __name__ == "__main__":
    main()
