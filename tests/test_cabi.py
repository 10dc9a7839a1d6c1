"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, exports every
symbol include/lm_b200.h declares, and its struct layouts match the ctypes mirrors.  No compute is launched."""
import ctypes
import os
import re

import pytest

from locomanipulationrl_b200 import _lib, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_library):
    lib = ctypes.CDLL(built_library)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lm_b200.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), set(declared) ^ set(_lib.SYMBOLS)


def test_loads_and_struct_layouts(built_library):
    lib = _lib.load()
    assert lib.lm_abi_version() == 1
    for which, struct in enumerate((_lib.LmTaskParams, _lib.LmRobotViews, _lib.LmTaskBuffers, _lib.LmHostStep, _lib.LmPdState,
                                    _lib.LmRunningStats, _lib.LmRolloutStep, _lib.LmRolloutBuffers, _lib.LmPpoBatch, _lib.LmNoiseSpec, _lib.LmNoiseState, _lib.LmCustomController, _lib.LmAdamSchedule, _lib.LmActionSink, _lib.LmHostPolicy)):
        assert lib.lm_struct_size(which) == ctypes.sizeof(struct)


def test_task_params_marshalling():
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    p = _lib.task_params(cfg, 4096)
    assert p.kind == configs.KIND_COTRAIN and p.num_envs == 4096 and p.num_states == 64 and p.zero_actions == 1
    assert abs(p.time_interval - 0.0083 * 4) < 1e-9 and p.joint1_mirrored == 0
    assert [p.dof_index[j] for j in range(12)] == list(range(12))
    assert abs(p.corners[0][1] + 0.115) < 1e-7 and abs(p.mani_default_position[2] - 0.6) < 1e-7
    q = _lib.task_params(configs.get_task_config("JointLocomanipulation"), 64)
    assert q.roll_range == 0.0 and abs(q.min_yaw - 0.785) < 1e-7          # JL ships fixed goal angles (JL:61-66)


def test_errors_without_gpu_are_loud(built_library):
    """No device here: creation must fail with a message, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    p = _lib.task_params(configs.get_task_config("QuadrupedPoseControl"), 128)
    assert not lib.lm_task_create(ctypes.byref(p), 0)
    assert b"CUDA" in lib.lm_last_error() or b"device" in lib.lm_last_error()
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    with pytest.raises(_lib.LmError):
        LocoManipulationTask(configs.get_task_config("QuadrupedPoseControl"), 128, device="cuda:0")
    with pytest.raises(_lib.LmError):
        LocoManipulationTask(configs.get_task_config("QuadrupedPoseControl"), 128, device="cpu")
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    with pytest.raises(_lib.LmError):
        MlpActorCritic(device="cpu").forward(torch.zeros(4, 64))
