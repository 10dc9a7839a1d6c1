"""Multi-GPU plumbing: one process per GPU, environments block-sharded, NCCL used for exactly two things —
the PPO gradient all-reduce and the episode-stat reduction (the reference only passes the
``multi_gpu`` flag through to rl_games, cfg/config.yaml:29, scripts/rlgames_train.py:93-96).

Environments never exchange data, so there is no data-path collective.  Each rank keeps its own
locomotion/manipulation halves (N_r/2 + N_r/2); ``EnvSharding.global_ids`` gives the fixed permutation back to
the reference's global ``[all loco ; all mani]`` order for parity checks.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> "tuple[int, int, int]":
    """Join the torchrun rendezvous if WORLD_SIZE > 1.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


class PeerComm:
    """One-shot all-reduce over NVLink peer memory (csrc/lm_comm.cuh): every rank pushes its values into every peer's receive
    slot, raises a flag, and sums the slots in rank order.  Receive memory is allocated by the library; its cudaIpc handle is
    exchanged here through the process group that already exists (torch.distributed is the plumbing, the data path is ours).
    ``all_reduce`` is CUDA-graph capturable; the gradient path uses the same communicator from inside the weight-gradient
    kernel (``ppo.MlpPpoGradients(comm=...)``)."""

    def __init__(self, max_floats: int, device=None, group=None):
        import ctypes as C
        from . import _lib
        self._lib_mod, self._lib = _lib, _lib.load()
        if dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.handle = self._lib.lm_comm_create(idx, self.rank, self.world, int(max_floats))
        if not self.handle:
            raise _lib.LmError(f"lm_comm_create failed: {self._lib.lm_last_error().decode()}")
        self.max_floats = int(max_floats)
        if self.world > 1:
            nb = self._lib.lm_comm_handle_bytes()
            mine = (C.c_uint8 * nb)()
            _lib.check(self._lib.lm_comm_get_handle(self.handle, mine), "lm_comm_get_handle")
            backend = dist.get_backend(group)
            t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.device if backend == "nccl" else "cpu")
            gathered = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(gathered, t, group=group)
            blob = b"".join(bytes(g.cpu().tolist()) for g in gathered)
            buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
            _lib.check(self._lib.lm_comm_open(self.handle, buf), "lm_comm_open")
            dist.barrier(group=group)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a contiguous fp32 CUDA tensor (<= max_floats elements)."""
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise self._lib_mod.LmError("PeerComm.all_reduce: contiguous fp32 CUDA tensor required")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._lib_mod.check(self._lib.lm_comm_allreduce(self.handle, t.data_ptr(), t.numel(), stream), "lm_comm_allreduce")
        return t

    def all_reduce_gather(self, t: torch.Tensor, off: int, f32_src: torch.Tensor, f32_scale: float, i64_src: torch.Tensor,
                          one_more: float) -> torch.Tensor:
        """``all_reduce`` whose elements [off, off + len(f32_src) + len(i64_src) + 1) are f32_src * f32_scale, i64_src (-> fp32) and
        the constant ``one_more`` instead of ``t``'s own: one launch packs and reduces."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._lib_mod.check(self._lib.lm_comm_allreduce_gather(self.handle, t.data_ptr(), t.numel(), int(off), f32_src.data_ptr(), f32_src.numel(),
                                                               float(f32_scale), i64_src.data_ptr(), i64_src.numel(), float(one_more), stream),
                            "lm_comm_allreduce_gather")
        return t

    def status(self) -> int:
        """0, or 1 + r after a time-out waiting for rank r (host synchronisation)."""
        return int(self._lib.lm_comm_status(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.lm_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass
class EnvSharding:
    """Block sharding of ``global_envs`` environments over ``world`` ranks."""
    global_envs: int
    world: int
    rank: int
    cotrain: bool = True

    def __post_init__(self):
        unit = 2 * self.world if self.cotrain else self.world
        if self.global_envs % unit:
            raise ValueError(f"global env count {self.global_envs} must be a multiple of {unit}")

    @property
    def local_envs(self) -> int:
        return self.global_envs // self.world

    def global_ids(self) -> torch.Tensor:
        """Global env id (reference order) of every local env of this rank."""
        n = self.local_envs
        if not self.cotrain:
            return torch.arange(n) + self.rank * n
        h_local, h_global = n // 2, self.global_envs // 2
        loco = torch.arange(h_local) + self.rank * h_local
        return torch.cat([loco, loco + h_global])

    def scatter_global(self, tensor_global: torch.Tensor) -> torch.Tensor:
        return tensor_global[self.global_ids().to(tensor_global.device)]


class StatReducer:
    """Sum-reduces the per-rank reward-term means (weighted by env count) and the success / reset counters, then applies the
    reference's success-rate bookkeeping on the global counts (QPC:619-633 / JLV:758-802).  One small all-reduce."""

    def __init__(self, num_local_envs: int, device, group=None):
        self.n = num_local_envs
        self.group = group
        self.buf = torch.zeros(7 + 6 + 1, dtype=torch.float64, device=device)

    def reduce(self, extras_buf: torch.Tensor, counters: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.buf[0:7] = extras_buf[0:7].double() * self.n
        self.buf[7:13] = counters.double()
        self.buf[13] = float(self.n)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        total = self.buf[13]
        return {"reward_means": (self.buf[0:7] / total).float(), "counters": self.buf[7:13].round().long(), "num_envs": total.long()}


class GradArena:
    """Pre-flattened fp32 gradient buffer: parameters' ``.grad`` are views into one arena, so the PPO gradient all-reduce
    (rl_games ``multi_gpu``: all_reduce of the concatenated grads / world) is ONE NCCL call with no ``torch.cat``.
    ``tail`` extra fp32 slots ride along in the same call (``FusedHorizonReducer`` puts the episode stats there: at these
    sizes a collective is pure latency, so one call instead of two halves the cost of the rollout boundary)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, tail: int = 0, comm: Optional[PeerComm] = None):
        self.comm = comm
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.storage = torch.zeros(n + tail, dtype=torch.float32, device=dev)
        self.flat = self.storage[:n]
        self.tail = self.storage[n:]
        self.group = group
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def numel(self) -> int:
        return self.flat.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def allreduce_mean(self, async_op: bool = False):
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return None
        if self.comm is not None:                     # one-shot NVLink all-reduce (our kernel) instead of the NCCL call
            self.comm.all_reduce(self.storage)
            self.flat.div_(self.comm.world)
            return None
        work = dist.all_reduce(self.storage, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(dist.get_world_size(self.group))
        return None


class FusedHorizonReducer:
    """Everything that crosses GPUs at a rollout boundary in ONE all-reduce: the gradient arena plus, in its 14-slot tail,
    [7 reward-term sums (mean x local envs), 6 success / reset counters, env count].  fp32 is exact for the counters
    (< 2^24) and costs the means one rounding (6e-8 relative).  With a ``PeerComm`` the whole boundary is ONE launch (the all-reduce
    kernel gathers the statistics into the message itself); through NCCL it is three small torch ops and the collective.  The
    gradient part holds the MEAN over ranks afterwards (NCCL path) or the SUM (``PeerComm`` path: ``FlatAdam.step(grad_scale=
    1 / world)`` averages inside its own launch) — ``grad_scale`` says which."""
    TAIL = 14

    def __init__(self, params: Iterable[torch.nn.Parameter], num_local_envs: int, group=None, comm: Optional[PeerComm] = None):
        self.arena = GradArena(params, group=group, tail=self.TAIL, comm=None)
        self.n = float(num_local_envs)
        self.comm, self.group = comm, group

    @property
    def grad_scale(self) -> float:
        """What ``FlatAdam.step(grad_scale=...)`` must apply to the arena after :meth:`reduce`."""
        return 1.0 / self.comm.world if (self.comm is not None and self.comm.world > 1) else 1.0

    def reduce(self, extras_buf: torch.Tensor, counters: torch.Tensor) -> None:
        a = self.arena
        if self.comm is not None and self.comm.world > 1:
            self.comm.all_reduce_gather(a.storage, a.numel, extras_buf[0:7], self.n, counters, self.n)
            return
        t = a.tail
        torch.mul(extras_buf[0:7], self.n, out=t[0:7])
        t[7:13].copy_(counters)
        t[13].fill_(self.n)
        a.allreduce_mean()

    def result(self) -> Dict[str, torch.Tensor]:
        t = self.arena.tail
        return {"reward_means": t[0:7] / t[13], "counters": t[7:13].round().long(), "num_envs": t[13].round().long()}
