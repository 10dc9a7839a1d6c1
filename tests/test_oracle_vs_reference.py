"""Live check of the oracle port against the UNMODIFIED reference classes (oracle/refshim.py) on fresh seeds.

Runs only where /root/reference exists (the authoring container); the GPU box checks the same thing through the committed
fixtures in tests/golden/ (oracle/make_golden.py).  Bit-exact: the port issues the same torch / SciPy calls."""
import os

import pytest
import torch

from locomanipulationrl_b200 import configs, synthetic

REF = "/root/reference/RobotLearning/omniisaacgymenvs"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this box")


def _inputs(n, steps, seed):
    g = torch.Generator().manual_seed(seed)
    return ([(torch.rand(n, 12, generator=g) * 2 - 1) for _ in range(steps)], [torch.rand(n, 3, generator=g) for _ in range(steps)])


@pytest.mark.parametrize("name,n,seed", [("JointLocomanipulationVertical", 512, 101), ("QuadrupedManipulatePlate", 256, 102),
                                         ("QuadrupedPoseControl", 256, 103), ("JointLocomanipulation", 128, 104),
                                         ("QuadrupedPoseControlVertical", 128, 105), ("QuadrupedManipulatePlateVertical", 128, 106)])
def test_port_is_bit_exact_to_the_reference_classes(name, n, seed):
    from oracle import port, refshim
    cfg = configs.get_task_config(name)
    steps = 22                                                     # past the 15 consecutive successes that trigger goal resets
    scene_ref = synthetic.SyntheticScene(cfg, n, seed=seed)
    scene_port = synthetic.SyntheticScene(cfg, n, seed=seed)
    ref = refshim.make_reference_task(cfg, scene_ref)
    p = port.TaskPort(cfg, scene_port)
    acts, unis = _inputs(n, steps, seed)
    for t in range(steps):
        with refshim.supplied_uniforms(unis[t]), refshim.quiet():
            ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1).clone()
            ref.pre_physics_step(acts[t].clone())
        p.pre_physics_step(acts[t].clone(), unis[t])
        assert torch.equal(p.reset_env_ids, ids), f"step {t}: reset ids"
        assert torch.equal(p.goal_quaternions, ref.goal_quaternions), f"step {t}: goals"
        for grp in ("loco", "mani"):
            for k, v in getattr(scene_ref, grp).items():
                assert torch.equal(getattr(scene_port, grp)[k], v), f"step {t}: scene.{grp}.{k} after reset"
        st = scene_ref.generate(t, ref.goal_quaternions)
        scene_ref.load(st); scene_port.load(st)
        with refshim.quiet():
            obs, rew, reset, extras = ref.post_physics_step()
        p.post_physics_step()
        assert torch.equal(p.obs_buf, obs), f"step {t}: obs"
        assert torch.equal(p.rew_buf, rew), f"step {t}: rew"
        assert torch.equal(p.reset_buf, reset.long()), f"step {t}: reset"
        for k in ("progress_buf", "successes", "consecutive_successes", "goal_reset_buf"):
            assert torch.equal(getattr(p, k), getattr(ref, k).long()), f"step {t}: {k}"
        assert torch.equal(p.last_base_tip_positions, ref.last_base_tip_positions)
        for k, v in extras.items():
            assert abs(float(p.extras[k]) - float(v)) <= 1e-7 * max(1.0, abs(float(v))), f"step {t}: extras {k}"
