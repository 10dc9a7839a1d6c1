"""locomanipulationrl_b200 — B200-native task / policy hot path of LocoManipulationRL behind the reference's task surface.

Python here is a thin host layer over liblm_b200.so (include/lm_b200.h): PyTorch owns tensors, the CUDA
library does the work, and nothing falls back to the CPU.
"""
from .configs import TaskConfig, get_task_config, TASK_CONFIGS  # noqa: F401

__all__ = ["TaskConfig", "get_task_config", "TASK_CONFIGS"]
