#!/usr/bin/env python
"""Benchmark of the task + policy hot path (BASELINE.json metric: env-steps/s, obs+reward+PD/targets+reset+policy).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU torch path (oracle port) on the host cores

Workload (config.workload): JointLocomanipulationVertical co-train (half locomotion, half manipulation envs) +
MLP actor-critic, 65,536 envs PER GPU — BASELINE.json configs[2] at N=1 and configs[4] (524,288 envs on 8 GPUs) at
N=8, i.e. weak scaling.  One "step" = pre_physics_step(actions) [reset scan, reset scatter + goal resample, action ->
targets] -> synthetic physics state swap -> post_physics_step [obs, reward, termination, stats] -> MLP forward.
Other workloads: --workload manip16k (configs[1]), gnn_gae (configs[3]).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same step driven through the
C-ABI host entry (lm_task_step_host) with pinned-host inputs/outputs copied every step.

Timed region: whole rollout horizons.  --steps K is rounded UP to a multiple of the horizon (48 for the MLP co-train workload,
the reference yaml; 16 for gnn_gae) and repeated until the region lasts >= 0.2 s; the line's `steps` is the number of steps
actually timed (`steps_requested` = K).  Every horizon ends with the rollout-boundary collective (episode statistics + the
gradient arena, one message), so a multi-rank line always has `collectives_in_timed_region` >= 1.  The `train` block is
the same loop with real work at the boundary: rollout writer every step, then GAE and the PPO update (5 mini-epochs x 8
minibatches, tcgen05 gradients, the gradient all-reduce fused into the weight-gradient reduction, clip + Adam).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ENVS_PER_GPU = 65536
STATE_SETS = 6            # distinct synthetic state sets rotated through: 6 x ~23 MB state + buffers > 126 MB L2
B_TASK = 1036             # algorithmic bytes / env-step, task path
B_POST = 892              # of which the post_physics_step kernel moves (DESIGN.md "k_post bytes")
MLP_FLOPS = 116352        # FLOP / env
HORIZON = 16              # GAE rollout length of the gnn_gae workload (BASELINE configs[3]: "16-step rollout")
HORIZON_MLP = 48          # rl_games horizon_length of the co-train yaml (cfg/train/JointLocomanipulationVerticalPPO.yaml:66)
MIN_TIMED_S = 0.2         # the timed region is at least this long whatever --steps says

WORKLOADS = {
    "cotrain64k": dict(task="JointLocomanipulationVertical", envs=ENVS_PER_GPU, policy="mlp",
                       desc="JointLocomanipulationVertical co-train + MLP actor-critic, 65,536 envs per GPU"),
    "manip16k": dict(task="QuadrupedManipulatePlate", envs=16384, policy="mlp",
                     desc="QuadrupedManipulatePlate + MLP actor-critic, 16,384 envs per GPU"),
    "gnn_gae": dict(task="JointLocomanipulationVertical", envs=ENVS_PER_GPU, policy="gnn",
                    desc="JointLocomanipulationVertical co-train + GNN policy + GAE over 16-step rollouts, 65,536 envs per GPU"),
}


def bench_config(w, gpus):
    """The `config` block, identical in both arms (the driver compares them key by key); run details go under `run`."""
    return {"workload": w["desc"], "envs_per_gpu": w["envs"], "global_envs": w["envs"] * gpus, "parallelism": f"env-shard x{gpus}",
            "l2": f"{STATE_SETS} rotating synthetic state sets (> L2) - inputs larger than L2"}


def measured_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` capture of this
    kernel at the bench workload (profiles/traffic.json; a profiler measurement, never taken inside a timed run)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f).get(kernel)
        return None if d is None else float(d["dram_read_bytes"]) + float(d["dram_write_bytes"])
    except (OSError, ValueError, KeyError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((int(s[1]) for s in self.samples if s[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- reference / CPU arm
def run_cpu_port(workload, steps, warmup, envs):
    """The reference's CPU torch path (oracle port: same scipy/torch calls as the reference classes) + torch policy."""
    from locomanipulationrl_b200 import configs, synthetic
    from oracle import port, port_policy
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOADS[workload]
    cfg = configs.get_task_config(w["task"])
    scene = synthetic.SyntheticScene(cfg, envs)
    task = port.TaskPort(cfg, scene)
    torch.manual_seed(0)
    if w["policy"] == "mlp":
        net = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ELU(), torch.nn.Linear(256, 128), torch.nn.ELU(),
                                  torch.nn.Linear(128, 64), torch.nn.ELU())
        mean, val = torch.nn.Linear(64, 12), torch.nn.Linear(64, 1)
        def policy(obs):
            h = net(obs)
            return mean(h), val(h)
    else:
        import numpy as np
        g = np.load(os.path.join(ROOT, "tests", "golden", "gnn_reference.npz"))
        sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p/")}
        def policy(obs):
            return port_policy.gnn_forward(sd, obs)
    states = [scene.generate(i) for i in range(STATE_SETS)]
    gen = torch.Generator().manual_seed(1)
    unis = [torch.rand(envs, 3, generator=gen) for _ in range(STATE_SETS)]
    actions = torch.zeros(envs, 12)
    times = []
    rollout = []
    with torch.no_grad():
        for t in range(warmup + steps):
            t0 = time.perf_counter()
            task.pre_physics_step(torch.clamp(actions, -1.0, 1.0), unis[t % STATE_SETS])
            scene.load(states[t % STATE_SETS])
            obs, rew, reset, _ = task.post_physics_step()
            mu, value = policy(torch.clamp(obs, -5.0, 5.0))
            actions = mu
            if w["policy"] == "gnn":
                rollout.append((rew.clone(), value.view(-1).clone(), reset.float()))
                if len(rollout) == HORIZON:
                    r, v, d = (torch.stack(x) for x in zip(*rollout))
                    port_policy.gae_rlgames(r, v, d, v[-1], d[-1])
                    rollout = []
            dt = time.perf_counter() - t0
            if t >= warmup:
                times.append(dt)
    total = sum(times)
    return envs * len(times) / total, total / len(times), torch.get_num_threads()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    envs = w["envs"]
    # bound the run: calibrate one step at full size, shrink the env sample if K steps would take too long
    v1, t1, cores = run_cpu_port(args.workload, 1, 1, envs)
    budget = 150.0
    while (args.steps + args.warmup) * t1 * (envs / w["envs"]) > budget and envs > 2048:
        envs //= 2
    value, per_step, cores = run_cpu_port(args.workload, args.steps, args.warmup, envs)
    sample = f"{args.steps} steps x {envs} envs of the {w['envs']}-env workload (oracle port: the reference's scipy/torch CPU path)"
    line = {"impl": "reference", "metric": "env-steps/s (obs+reward+PD+reset+policy)", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(w, args.gpus), "run": {"reference_sample_envs": envs},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- CUDA arm
def bind_host_cores(local_rank, world):
    """One disjoint slice of the host cores per rank (launch thread + pinned-buffer first touch stay put); NUMA-local to the rank's
    GPU when the platform exposes more than one node.  A single rank is bound to its GPU's local cores as a whole (on the 8-GPU
    boxes an unbound single rank measured 0.92 ms per host-buffer step against 0.67-0.74 for the bound ranks of a 2-rank run)."""
    if not hasattr(os, "sched_setaffinity"):
        return
    if os.environ.get("LM_BENCH_CORES"):              # experiment override, e.g. "0-7" or "4,5,6"
        cs = []
        for part in os.environ["LM_BENCH_CORES"].split(","):
            lo, _, hi = part.partition("-")
            cs += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cs)
        return
    try:
        cores = sorted(os.sched_getaffinity(0))
        node_cores = None
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/local_cpulist"
            if os.path.exists(path):
                node_cores = []
                for part in open(path).read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    node_cores += list(range(int(lo), int(hi or lo) + 1))
                node_cores = [c for c in node_cores if c in cores]
        except Exception:
            node_cores = None
        pool = node_cores if node_cores and len(node_cores) >= world else cores
        per = max(1, len(pool) // world)
        mine = pool[local_rank * per:(local_rank + 1) * per] or pool
        if world == 1 and node_cores is None:
            return                                  # nothing known about locality: leave the process where it is
        os.sched_setaffinity(0, mine)
    except Exception as exc:                      # affinity is an optimisation, never a failure
        print(f"bench.py: core binding skipped ({exc})", file=sys.stderr)


class Pipeline:
    """Device-resident step loop: S rotating state sets, post + policy + pre = 3 launches per step, captured as one CUDA graph per
    rollout horizon (see capture())."""

    def __init__(self, workload, device, rank):
        from locomanipulationrl_b200 import _lib, configs, synthetic
        from locomanipulationrl_b200.tasks.base import LocoManipulationTask
        from locomanipulationrl_b200.policy.mlp import MlpActorCritic
        from locomanipulationrl_b200.policy.gnn import GnnActorCritic
        self.w = WORKLOADS[workload]
        self.device = device
        self.n = n = self.w["envs"]
        self.lib = _lib.load()
        self._lib_mod = _lib
        cfg = configs.get_task_config(self.w["task"])
        self.cfg = cfg
        self.scenes = [synthetic.SyntheticScene(cfg, n, device=device, seed=1000 * rank + 42, env_id_offset=rank * n) for _ in range(STATE_SETS)]
        self.host_states = []
        for i, sc in enumerate(self.scenes):
            st = sc.generate(i)
            sc.load(st)
            self.host_states.append(sc.pack_host(st))
        self.task = LocoManipulationTask(cfg, n, scene=self.scenes[0], device=device, fuse_obs_clamp=True, seed=rank)
        torch.manual_seed(0)          # identical weights on every rank (data parallel)
        if self.w["policy"] == "mlp":
            self.policy = MlpActorCritic(device=device)
        else:
            self.policy = GnnActorCritic(device=device, precision=os.environ.get("LM_GNN_PRECISION", "bf16"))
        self.policy.pack()
        g = torch.Generator(device=device).manual_seed(7 + rank)
        self.uniforms = [torch.rand(n, 3, device=device, generator=g) for _ in range(STATE_SETS)]
        self.mu = torch.zeros(n, 12, device=device)
        self.value = torch.zeros(n, 1, device=device)
        self.norm_mean = torch.zeros(64, device=device)
        self.norm_inv_std = torch.ones(64, device=device)
        self.graphs = None
        self.gae_buf = None
        self.actions = self.mu                          # MLP workload: the policy mean is the next action
        # kernels of liblm_b200.so per step: k_pre, k_post_lean, policy forward (k_reset_count only runs when reset_buf was
        # written from outside); the gnn_gae workload adds k_rollout_write every step and lm_gae once per HORIZON steps
        self.launches_per_step = 3
        # Default chain (MLP workloads): post -> policy -> reset half of the next pre, where
        #   * the action half of pre_physics_step (adapter clamp, current_actions, take_action) is done by the MLP head epilogue
        #     (LmActionSink; LM_BENCH_NO_SINK=1: the actions go through k_pre),
        #   * the reset half, which needs nothing from the policy kernel, is released beside it by programmatic dependent launch and
        #     only completes after it (lm_task_set_early_reset; LM_BENCH_NO_EARLY=1: it runs behind the policy kernel),
        #   * the post kernel's finish pass (extras, counters) is one extra CTA of that launch (lm_task_set_defer_finish;
        #     LM_BENCH_INLINE_FINISH=1: the last-CTA tail inside the post kernel, the library's default mode).
        # Measured and dropped in round 2: the reset half on a second stream with an event join (+0.3 us), the reset half
        # in front of the policy kernel on the same stream (+1.6 us), the finish pass as a kernel on a second stream (+4.5 us).
        self.use_sink = self.w["policy"] == "mlp" and not os.environ.get("LM_BENCH_NO_SINK")
        self.early_reset = self.use_sink and not os.environ.get("LM_BENCH_NO_EARLY")
        self.graph_parity = None
        self.defer_finish = not bool(os.environ.get("LM_BENCH_INLINE_FINISH"))
        if self.defer_finish:
            self.task.set_defer_finish(True)
        if self.early_reset:
            self.task.set_early_reset(True)
            # step order post(t) -> policy(t) -> pre(t+1), so that the policy kernel and the reset half it overlaps sit in the same
            # captured graph; prime with the first step's pre_physics_step
            self.task.pre_physics_step(self.actions, self.uniforms[0])
        if self.w["policy"] == "gnn":
            from locomanipulationrl_b200.rollout import RolloutBuffer, RunningMeanStd
            T = HORIZON
            self.rollout = RolloutBuffer(T, n, device=device, reward_scale=0.01)
            self.obs_stats = RunningMeanStd(64, device=device)
            self.log_std = self.policy.log_std_parameter.detach()
            self.actions = torch.zeros(n, 12, device=device)
            self.gae_buf = dict(adv=torch.zeros(T, n, device=device), ret=torch.zeros(T, n, device=device))
            self.launches_per_step = 4

    def step_eager(self, i, t=0):
        task = self.task
        if self.early_reset:
            task.set_up_scene(self.scenes[i])           # "physics": the next state set becomes the live views
            task.post_physics_step()
            self.policy(task.obs_buf, norm_mean=self.norm_mean, norm_inv_std=self.norm_inv_std, out=(self.mu, self.value), sink=task.action_sink())
            task.pre_physics_step(None, self.uniforms[(i + 1) % STATE_SETS])     # reset half, released beside the policy kernel
            return
        # action sink (default): the kernel that produced the actions (MLP head epilogue / rollout writer) already delivered them into
        # the task's buffers (clamp, current_actions, joint targets), so pre_physics_step runs its reset half only
        task.pre_physics_step(None if self.use_sink else self.actions, self.uniforms[i])
        task.set_up_scene(self.scenes[i])               # "physics": the next state set becomes the live views
        task.post_physics_step()
        if self.w["policy"] == "mlp":
            self.policy(task.obs_buf, norm_mean=self.norm_mean, norm_inv_std=self.norm_inv_std, out=(self.mu, self.value),
                        sink=task.action_sink() if self.use_sink else None)
        else:
            self.policy(task.obs_buf, out=(self.mu, self.value))
            k = t % HORIZON
            self.rollout_step(k)
            if k == HORIZON - 1:
                b, td = self.gae_buf, self.rollout.tensor_dict
                s = torch.cuda.current_stream().cuda_stream
                self._lib_mod.check(self.lib.lm_gae(td["rewards"].data_ptr(), td["values"].data_ptr(), td["dones"].data_ptr(), td["values"][-1].data_ptr(),
                                                    td["dones"][-1].data_ptr(), b["adv"].data_ptr(), b["ret"].data_ptr(), HORIZON, self.n,
                                                    0.99, 0.95, 0, s), "lm_gae")

    def rollout_step(self, k):
        """Slot k of the time-major rollout <- this step (obs, mean action, value, shaped reward, done, neglogp), observation
        statistics folded in, actions handed to the next pre_physics_step: one launch (lm_rollout_write)."""
        task = self.task
        self.rollout.write(k, task.obs_buf, self.mu, self.value, self.log_std, task.rew_buf, task.reset_buf, noise=None,
                           actions_out=self.actions, obs_stats=self.obs_stats, sink=task.action_sink() if self.use_sink else None)

    def capture(self, after_step=None, steps_per_graph=1):
        """CUDA graphs of the step loop.  ``steps_per_graph`` = 1: one graph per state set (parity of the action ring alternates,
        STATE_SETS is even); = STATE_SETS: ONE graph holding a whole rotation of the state sets, replayed by the first step of every
        rotation (bench.py's default: the programmatic-dependent-launch chain then also spans the step boundaries, which a graph launch
        per step cuts).  ``after_step(i)``: called after each eager warm-up step (tests/test_bench_shapes_gpu.py tracks the oracle
        through them)."""
        assert STATE_SETS % 2 == 0 and (steps_per_graph == 1 or steps_per_graph % STATE_SETS == 0)
        if self.w["policy"] != "mlp":
            return                                      # the GAE cadence needs a host-side counter: stay eager
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(STATE_SETS):
                self.step_eager(i)                      # warm everything on the capture stream
                if after_step is not None:
                    after_step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.graph_parity, self.steps_per_graph = [], [], steps_per_graph
        for i in range(0, STATE_SETS, steps_per_graph):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for j in range(steps_per_graph):
                    self.step_eager((i + j) % STATE_SETS)
            self.graphs.append(g)
            self.graph_parity.append(self.lib.lm_task_step_parity(self.task._handle))
        torch.cuda.synchronize()

    def step(self, t):
        """Step t of the loop.  With a whole rotation per graph the first step of a rotation launches all of it (callers step in whole
        rotations: the rollout horizons are multiples of STATE_SETS)."""
        if self.graphs is not None:
            spg = self.steps_per_graph
            if t % spg == 0:
                k = (t % STATE_SETS) // spg if spg <= STATE_SETS else 0
                self.graphs[k].replay()
                # replays skip the host side of pre_physics_step: keep the handle's ring parity in step for host-side readers
                self.lib.lm_task_set_step_parity(self.task._handle, self.graph_parity[k])
        else:
            self.step_eager(t % STATE_SETS, t)


def time_kernel(fn, iters):
    """Average device time of one call sequence (CUDA events on the launching stream)."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(ev):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    return sum(ts) / len(ts)


def cuda_arm(args):
    from locomanipulationrl_b200 import dist as lmdist
    import torch.distributed as tdist
    rank, world, local_rank = lmdist.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU baseline)")
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    # The CUDA arm needs no CPU parallelism, and an OpenMP pool left spinning behind the synthetic-state generation steals the cores
    # the launch thread and the driver's copy threads run on: measured 0.72 -> 0.66 ms per host-buffer step on a 24-core box (what
    # torchrun's OMP_NUM_THREADS=1 did for the multi-rank runs).  The CPU baseline sets its own thread count (run_cpu_port).
    torch.set_num_threads(1)
    bind_host_cores(local_rank, world)
    pipe = Pipeline(args.workload, device, rank)
    n = pipe.n
    # rollout-boundary collective: episode statistics + the gradient arena in ONE message.  Default: our one-shot NVLink
    # all-reduce (lm_comm: peer stores + flag + local sum); --collective nccl: the same message through torch.distributed.
    peer = None
    if world > 1 and args.collective == "peer":
        try:
            peer = lmdist.PeerComm(max_floats=65536, device=device)
        except Exception as exc:
            print(f"bench.py: peer-memory communicator unavailable ({exc}); using NCCL for the collective", file=sys.stderr)
            peer = None
    flags = torch.tensor([1.0 if (peer is not None or world == 1) else 0.0], device=device)
    if world > 1:
        tdist.all_reduce(flags, op=tdist.ReduceOp.MIN)             # all ranks or none
        if float(flags.item()) == 0.0:
            peer = None
    reducer = lmdist.FusedHorizonReducer(pipe.policy.parameters(), n, comm=peer)
    horizon = HORIZON_MLP if pipe.w["policy"] == "mlp" else HORIZON
    comm_stream = torch.cuda.Stream(device=device)

    comm_graph = None

    def comm_body():
        reducer.reduce(pipe.task._extras_buf, pipe.task._counters)

    def comm():
        """Episode-stat reduction + gradient arena, once per rollout horizon (one CUDA-graph replay: the host must not fall behind a
        45 us step)."""
        if world == 1 or os.environ.get("LM_BENCH_NO_COMM"):
            return 0
        if args.comm_stream == "main":
            # at the rollout boundary, on the launching stream (where the PPO update that consumes the gradients sits).
            # On a side stream the collective's CTAs pin SMs while they wait for the peer rank, and the persistent MLP kernel
            # (one 219 KB CTA per SM) then stalls on exactly those SMs: measured 216 us per event vs ~30 us in-line.
            if comm_graph is not None:
                comm_graph.replay()
            else:
                comm_body()
            return 1
        comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm_stream):
            if comm_graph is not None:
                comm_graph.replay()
            else:
                comm_body()
        return 1

    if world > 1:
        comm(); comm()                       # eager first: NCCL connects lazily at the first collective of each size
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=comm_stream if args.comm_stream == "side" else None):
                comm_body()
            comm_graph = g
        except Exception as exc:             # capture of the collective unsupported in this build: stay eager
            print(f"bench.py: comm graph capture failed ({exc}); using eager collectives", file=sys.stderr)
            comm_graph = None
        torch.cuda.synchronize()

    if not args.no_graphs:
        assert HORIZON_MLP % STATE_SETS == 0
        # one graph per rollout horizon (48 steps = 8 rotations of the state sets); LM_BENCH_GRAPH_STEPS: 1 or a multiple of STATE_SETS
        spg = int(os.environ.get("LM_BENCH_GRAPH_STEPS", "1" if os.environ.get("LM_BENCH_GRAPH_PER_STEP") else str(HORIZON_MLP)))
        assert HORIZON_MLP % spg == 0
        pipe.capture(steps_per_graph=spg)
    wq = getattr(pipe, "steps_per_graph", 1) if pipe.graphs is not None else 1
    wq = max(wq, STATE_SETS)
    warm_steps = -(-max(args.warmup, 3) // wq) * wq                        # whole graphs / rotations of the state sets
    for t in range(warm_steps):
        pipe.step(t)
    for _ in range(2):      # NCCL connects lazily at the first collective of each size: keep that out of the timed region
        comm()
    torch.cuda.current_stream().wait_stream(comm_stream)
    torch.cuda.synchronize()

    def run_horizons(reps):
        c = 0
        for r in range(reps):
            for k in range(horizon):
                pipe.step(r * horizon + k)
            if pipe.defer_finish:
                pipe.task.post_finish()       # extras / counters of the horizon's last step, ahead of the statistics reduction
            c += comm()
        return c

    # calibrate: whole horizons, >= the requested step count, >= MIN_TIMED_S of device time (every rank takes the same count)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); run_horizons(1); c1.record()
    torch.cuda.synchronize()
    per_h = torch.tensor([c0.elapsed_time(c1) * 1e-3], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(per_h, op=tdist.ReduceOp.MIN)
    reps = max((args.steps + horizon - 1) // horizon, int(MIN_TIMED_S / max(float(per_h.item()), 1e-6)) + 1)
    steps_timed = reps * horizon
    if world > 1:
        tdist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    collectives = run_horizons(reps)
    torch.cuda.current_stream().wait_stream(comm_stream)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    ms_total = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    t_ms = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(t_ms, op=tdist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    value = n * world * steps_timed / (ms_total * 1e-3)
    if world > 1 and collectives == 0 and not os.environ.get("LM_BENCH_NO_COMM"):
        raise SystemExit("bench.py: a multi-rank line needs at least one collective inside the timed region")
    comm_us = None
    if world > 1 and collectives > 0:
        # the collective alone (same graph, back to back on the launching stream), max over ranks: what a horizon boundary costs
        torch.cuda.synchronize(); tdist.barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(50):
            comm()
        torch.cuda.current_stream().wait_stream(comm_stream)
        k1.record(); torch.cuda.synchronize()
        t_c = torch.tensor([k0.elapsed_time(k1) / 50.0 * 1e3], device=device, dtype=torch.float64)
        tdist.all_reduce(t_c, op=tdist.ReduceOp.MAX)
        comm_us = float(t_c.item())
    if peer is not None and peer.status() != 0:
        raise SystemExit(f"bench.py: peer communicator reported a time-out (status {peer.status()})")

    # ---- per-kernel timing (rank 0; eager launches bracketed by events, same state rotation)
    line_extra = {}
    if rank == 0 and not args.no_kernel_times:
        task = pipe.task
        iters = max(12, min(args.steps, 60))
        if pipe.early_reset:
            task.set_early_reset(False)       # the sequences timed below put pre right behind post: outside the early-reset contract

        def post_only(i):
            task.set_up_scene(pipe.scenes[i % STATE_SETS])
            task._posted_since_pre = False     # time the kernel alone (no double-post fix-up copy)
            task.post_physics_step()

        def post_and_finish(i):
            post_only(i)

        def pre_only(i):
            task.pre_physics_step(None if pipe.use_sink else pipe.actions, pipe.uniforms[i % STATE_SETS])

        def policy_only(i):
            if pipe.w["policy"] == "mlp":
                pipe.policy(task.obs_buf, norm_mean=pipe.norm_mean, norm_inv_std=pipe.norm_inv_std, out=(pipe.mu, pipe.value),
                            sink=task.action_sink() if pipe.use_sink else None)
            else:
                pipe.policy(task.obs_buf, out=(pipe.mu, pipe.value))

        for i in range(3):
            pre_only(i); post_only(i); policy_only(i)
        torch.cuda.synchronize()

        def graph_time(fn, reps=20):
            """Average device time of one call: STATE_SETS calls captured back-to-back in a CUDA graph (no host launch gaps),
            replayed `reps` times between two CUDA events on the launching stream."""
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(STATE_SETS):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                g.replay()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / (reps * STATE_SETS)

        def pre_post(i):
            pre_only(i); post_and_finish(i)

        ms_post = graph_time(post_only)
        ms_post_fin = ms_post      # deferred finish: its CTA rides inside the pre launch and is timed with it
        ms_pre = max(graph_time(pre_post) - ms_post_fin, 1e-6)      # pre needs a preceding post + finish (its reset scan is their by-product)
        ms_pol = graph_time(policy_only)
        hbm, tf, src = measured_peaks()
        # single-task files with the 93-wide states_buf write 372 B/env more (1 036 -> 1 408 B/env-step)
        b_states = 372 if pipe.cfg.num_states == 93 else 0
        b_post, b_task = B_POST + b_states, B_TASK + b_states
        ach = n * b_post / (ms_post * 1e-3) / 1e9
        ach_task = n * b_task / ((ms_post + ms_pre) * 1e-3) / 1e9
        line_extra["roofline_post"] = {"kernel": "k_post (post_physics_step)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                                  "frac": ach / hbm, "traffic": measured_traffic("k_post_lean") if pipe.w["task"].startswith("Joint") else None,
                                  "traffic_unit": "bytes per launch (ncu dram read + write, 65,536 envs; stores mostly still in L2 at kernel end)",
                                  "algorithmic_bytes": n * b_post, "peak_source": src, "bytes_per_env": b_post, "ms": ms_post}
        line_extra["roofline_task"] = {"kernels": "k_reset_count + k_pre + k_post", "bound": "hbm", "achieved": ach_task, "peak": hbm,
                                       "unit": "GB/s", "frac": ach_task / hbm, "bytes_per_env": b_task, "ms": ms_post + ms_pre}
        if pipe.w["policy"] == "mlp":
            tfl = n * MLP_FLOPS / (ms_pol * 1e-3) / 1e12
            line_extra["roofline_policy"] = {"kernel": "k_mlp_forward (tcgen05)", "bound": "tensor", "achieved": tfl, "peak": tf, "unit": "TFLOP/s",
                                             "frac": tfl / tf, "ms": ms_pol, "hbm_gbs": n * 308 / (ms_pol * 1e-3) / 1e9}
        else:
            tc = pipe.policy.precision == "bf16"
            # issued tensor-core FLOP: 24 edges x (K = 64 then K = 32, N = 32) x 3 layers (the input layers and heads are fp32 CUDA-core work)
            gnn_flops = 24 * 2 * (64 * 32 + 32 * 32) * 3 if tc else 447296
            tfl = n * gnn_flops / (ms_pol * 1e-3) / 1e12
            line_extra["roofline_policy"] = {"kernel": "k_gnn_forward_tc (tcgen05, bf16)" if tc else "k_gnn_forward (fp32 CUDA cores)",
                                             "bound": "epilogue (TMEM->RF 64 B/clk/SM, MUFU ex2)" if tc else "fp32", "achieved": tfl,
                                             "peak": tf if tc else None, "frac": tfl / tf if tc else None,
                                             "unit": "TFLOP/s" if tc else "TFLOP/s (naive FLOP count)", "ms": ms_pol,
                                             "hbm_gbs": n * 308 / (ms_pol * 1e-3) / 1e9}
        line_extra["kernel_ms"] = {"pre": ms_pre, "post": ms_post, "policy": ms_pol, "post_finish": ms_post_fin - ms_post}
        # `roofline` = the dominant kernel of the step (largest share of the device time), the others keep their own blocks
        if ms_pol >= ms_post and line_extra["roofline_policy"].get("frac") is not None:
            line_extra["roofline"] = dict(line_extra["roofline_policy"], traffic=measured_traffic("k_mlp_forward") if pipe.w["policy"] == "mlp" else measured_traffic("k_gnn_forward_tc"),
                                          algorithmic_flops=n * (MLP_FLOPS if pipe.w["policy"] == "mlp" else 24 * 2 * (64 * 32 + 32 * 32) * 3),
                                          algorithmic_bytes=n * 308, peak_source=src,
                                          share_of_step=ms_pol / (ms_pre + ms_post + ms_pol))
        else:
            line_extra["roofline"] = dict(line_extra["roofline_post"], share_of_step=ms_post / (ms_pre + ms_post + ms_pol))
        if pipe.w["policy"] == "gnn":
            ms_ro = graph_time(lambda i: pipe.rollout_step(i % HORIZON))
            B_RO = 368 + 368 + 48        # obs/mu/value/rew/done in, obses/actions/mus/values/rewards/dones/neglogp out, + actions_out
            line_extra["roofline_rollout"] = {"kernel": "k_rollout_write", "bound": "hbm", "achieved": n * B_RO / (ms_ro * 1e-3) / 1e9, "peak": hbm,
                                              "unit": "GB/s", "frac": n * B_RO / (ms_ro * 1e-3) / 1e9 / hbm, "bytes_per_env": B_RO, "ms": ms_ro}
            line_extra["kernel_ms"]["rollout_write"] = ms_ro

            # 8 rotating rollouts (8 x 21 MB > the 126 MB L2), so the figure is HBM traffic, not L2 hits
            gsets = [tuple(torch.randn(HORIZON, n, device=device) for _ in range(3)) + (torch.empty(HORIZON, n, device=device), torch.empty(HORIZON, n, device=device))
                     for _ in range(8)]

            def gae_only(i):
                r_, v_, d_, adv_, ret_ = gsets[i % 8]
                pipe._lib_mod.check(pipe.lib.lm_gae(r_.data_ptr(), v_.data_ptr(), d_.data_ptr(), v_[-1].data_ptr(), d_[-1].data_ptr(), adv_.data_ptr(),
                                                    ret_.data_ptr(), HORIZON, n, 0.99, 0.95, 0, torch.cuda.current_stream().cuda_stream), "lm_gae")
            for i in range(8):
                gae_only(i)
            torch.cuda.synchronize()
            g8 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g8):
                for i in range(8):
                    gae_only(i)
            g8.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g8.replay()
            e1.record(); torch.cuda.synchronize()
            ms_gae = e0.elapsed_time(e1) / (20 * 8)
            b_gae = 20 * HORIZON + 8             # r, V, d in, advantages, returns out per (env, t) + last value / done per env
            line_extra["roofline_gae"] = {"kernel": "k_gae_time_major (once per 16-step rollout)", "bound": "hbm", "achieved": n * b_gae / (ms_gae * 1e-3) / 1e9,
                                          "peak": hbm, "unit": "GB/s", "frac": n * b_gae / (ms_gae * 1e-3) / 1e9 / hbm, "bytes_per_env": b_gae, "ms": ms_gae,
                                          "l2": "8 rotating rollouts (172 MB > L2)"}
            line_extra["kernel_ms"]["gae_per_rollout"] = ms_gae
        if world == 1 and not args.no_large:
            line_extra["roofline_large"] = large_roofline(device, graph_time, hbm)

    # ---- e2e: C-ABI host entry, pinned host buffers in, results out, every step
    e2e = None if args.no_e2e else e2e_arm(pipe, args, world, device, rank)
    if e2e is not None and world == 1:        # the same step with the observations copied back as well (a host-side policy's data flow)
        full = e2e_arm(pipe, args, world, device, rank, readback_obs=True, shards=4)     # 21 MB back per step: pipelining over shards pays here
        e2e["with_obs_readback"] = {k: full[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "api")}

    # ---- train block: the same loop with the rollout writer, GAE and the PPO update (real gradients in the all-reduce)
    train = None
    if pipe.w["policy"] == "mlp" and args.workload == "cotrain64k" and not args.no_train:
        train = train_arm(pipe, args, world, device, rank, peer)

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        w = WORKLOADS[args.workload]
        cpu_steps = 96 if w["policy"] == "mlp" else 48     # ~12 s (MLP workload) / ~3 s of CPU work on a 16-core host
        cpu_envs = w["envs"] if w["policy"] == "mlp" else 8192
        v, per_step, cores = run_cpu_port(args.workload, cpu_steps, 2, cpu_envs)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_steps} steps x {cpu_envs} envs of the same workload (oracle port = the reference's scipy/torch CPU path), {per_step * 1e3:.0f} ms/step"}
    if rank == 0:
        clocks = sampler.summary()
        line = {"metric": "env-steps/s (obs+reward+PD+reset+policy)", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": steps_timed, "steps_requested": args.steps, "warmup": warm_steps, "ms_per_step": ms_total / steps_timed,
                "timed_region_s": ms_total * 1e-3, "collectives_in_timed_region": collectives, "collective_us": comm_us,
                "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 task math, bf16 x bf16 -> f32 policy GEMMs",
                "data": "synthetic",
                "config": bench_config(pipe.w, world),
                "run": {"cuda_graphs": pipe.graphs is not None, "action_sink": pipe.use_sink,
                           "post_finish_deferred": pipe.defer_finish, "early_reset": pipe.early_reset, "comm": (f"one stat + gradient-arena all-reduce (58,651 fp32) every {horizon} steps, "
                                    + ("one-shot NVLink peer-memory kernel (lm_comm_allreduce)" if peer is not None else "NCCL all_reduce")
                                    + f", {args.comm_stream} stream, CUDA graph") if world > 1 else "none",
                           "timed": f"{reps} rollout horizons of {horizon} steps"},
                "gpu_launches": pipe.launches_per_step * steps_timed + (steps_timed // HORIZON if pipe.w["policy"] == "gnn" else 0)
                                + (collectives if peer is not None else 0) + (reps if pipe.defer_finish else 0), "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu, "train": train}
        line.update(line_extra)
        emit(line)
    if world > 1:
        # The captured NCCL graph keeps communicator resources alive and destroy_process_group() then blocks (observed: the
        # process sat in teardown until an outer timeout).  Everything is flushed and every rank has passed the barrier:
        # leave without the teardown.
        torch.cuda.synchronize()
        tdist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def train_arm(pipe, args, world, device, rank, peer):
    """The same env loop with real work at the rollout boundary (cfg/train/JointLocomanipulationVerticalPPO.yaml: horizon 48, 5
    mini-epochs, minibatch = batch / 8, normalize_input / normalize_value, adaptive schedule at kl_threshold 0.008):
    per step pre_physics_step -> state swap -> post_physics_step -> MLP forward on the normalised observations -> rollout write
    (in-kernel action sampling); per rollout finish -> GAE -> PpoUpdater.update, whose per-minibatch gradient all-reduce is fused
    into the weight-gradient reduction over NVLink peer memory (N > 1).  Device time, max over ranks."""
    import torch.distributed as tdist
    from locomanipulationrl_b200.ppo import PpoUpdater
    from locomanipulationrl_b200.rollout import RolloutBuffer, RunningMeanStd
    T, n, task, policy = HORIZON_MLP, pipe.n, pipe.task, pipe.policy
    if pipe.defer_finish:
        task.post_finish()
        task.set_defer_finish(False)
    task.set_up_scene(pipe.scenes[0])                  # whatever pipeline variant ran before: start from a plain post step
    task.post_physics_step()
    buf = RolloutBuffer(T, n, device=device, reward_scale=1.0)
    obs_rms = RunningMeanStd(64, device=device)
    nmb = 8
    up = PpoUpdater(policy, buf, minibatch_size=T * n // nmb, mini_epochs=5, normalize_value=True, obs_rms=obs_rms, comm=peer)
    log_std = policy.log_std_parameter.detach()
    actions = torch.zeros(n, 12, device=device)
    mu, value = torch.zeros(n, 12, device=device), torch.zeros(n, 1, device=device)
    state = {"i": 0}

    def forward():
        mean, inv = obs_rms.export()
        policy(task.obs_buf, norm_mean=mean, norm_inv_std=inv, out=(mu, value))

    def iteration(seed):
        for t in range(T):
            forward()
            buf.write(t, task.obs_buf, mu, value, log_std, task.rew_buf, task.reset_buf, sample_seed=seed, value_denorm=up.value_denorm,
                      actions_out=actions)
            i = state["i"] = (state["i"] + 1) % STATE_SETS
            task.pre_physics_step(actions, pipe.uniforms[i])
            task.set_up_scene(pipe.scenes[i])
            task.post_physics_step()
        forward()                                         # value of the observation after the last stored action: GAE's bootstrap
        lv, ld = buf.finish(task.rew_buf, value, task.reset_buf, value_denorm=up.value_denorm)
        return up.update(lv, ld)

    iteration(1)
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    iters = 2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for k in range(iters):
        stats = iteration(2 + k)
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / iters], device=device, dtype=torch.float64)
    # the update alone (same rollout again): what the boundary costs
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        tdist.barrier()
    u0.record()
    up.update(buf.last_values, buf.last_dones)
    u1.record(); torch.cuda.synchronize()
    ms_up = torch.tensor([u0.elapsed_time(u1)], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(ms_up, op=tdist.ReduceOp.MAX)
    st = stats.detach().float().cpu() / world
    if peer is not None and peer.status() != 0:
        raise SystemExit(f"bench.py: peer communicator reported a time-out (status {peer.status()})")
    return {"metric": "env-steps/s incl. rollout write, GAE and the PPO update", "value": n * world * T / (float(ms.item()) * 1e-3), "unit": "env-steps/s",
            "ms_per_iteration": float(ms.item()), "ms_update": float(ms_up.item()), "ms_rollout": float(ms.item()) - float(ms_up.item()),
            "horizon": T, "mini_epochs": 5, "minibatches": nmb, "minibatch_size": T * n // nmb, "samples_per_update": T * n * 5,
            "collectives_per_iteration": (5 * nmb if world > 1 else 0),
            "allreduce": ("fused into k_wgrad_reduce_allreduce: 58,645 fp32 per minibatch pushed over NVLink peer memory" if peer is not None
                          else ("torch.distributed all_reduce of the flat arena per minibatch" if world > 1 else "none")),
            "loss_stats_last_minibatch": {k: float(st[i]) for i, k in enumerate(("a_loss", "c_loss", "b_loss", "entropy", "kl", "loss", "clip_frac"))},
            "lr_after": up.opt.lr}


def large_roofline(device, graph_time, hbm, n=524288):
    """k_post at 524,288 envs on ONE GPU (BASELINE configs[4]'s global batch): several waves of CTAs, so the figure shows
    the kernel's steady-state quality rather than the launch/latency floor that dominates at 65,536 envs."""
    from locomanipulationrl_b200 import configs, synthetic
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    scenes = [synthetic.SyntheticScene(cfg, n, device=device, seed=77) for _ in range(2)]      # 2 x ~190 MB of state > L2
    small = synthetic.SyntheticScene(cfg, 65536, seed=77)
    st = small.generate(0)
    for sc in scenes:                          # tile the 65,536-env synthetic state 8x (generation cost, not content, matters here)
        for grp, dst in (("loco", sc.loco), ("mani", sc.mani)):
            for k, v in st[grp].items():
                reps = dst[k].shape[0] // v.shape[0]
                dst[k].copy_(v.repeat(reps, 1))
    task = LocoManipulationTask(cfg, n, scene=scenes[0], device=device, fuse_obs_clamp=True)
    task.pre_physics_step(torch.zeros(n, 12, device=device), torch.rand(n, 3, device=device))

    def post_only(i):
        task.set_up_scene(scenes[i % 2])
        task._posted_since_pre = False
        task.post_physics_step()

    for i in range(3):
        post_only(i)
    ms = graph_time(post_only, reps=10)
    ach = n * B_POST / (ms * 1e-3) / 1e9
    return {"kernel": "k_post_lean", "envs": n, "ms": ms, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "bytes_per_env": B_POST}


def e2e_arm(pipe, args, world, device, rank=0, readback_obs=False, shards=None):
    """Same step through the host-buffer API (ShardedHostStepper -> lm_task_step_host_async per env shard): H2D of
    actions / uniforms / physics state from pinned memory, kernels, D2H of rew / reset, the policy forward and D2H of its
    outputs (the next actions, value), every step, with a host synchronisation at the end of every step.  Max over ranks.
    The policy is device-resident, so by default the observations stay on the device (the reference's adapter hands them to
    rl_device = cuda too); ``readback_obs`` adds their D2H copy, the data flow of a host-side policy."""
    import torch.distributed as tdist
    from locomanipulationrl_b200.host_step import ShardedHostStepper
    n = pipe.n
    mlp = pipe.w["policy"] == "mlp"
    shards = shards or args.e2e_shards
    sizes = None
    if os.environ.get("LM_BENCH_E2E_SPLIT") and shards == 2:          # e.g. "0.75": the first shard's share of the envs (multiples of 256)
        first = int(round(n * float(os.environ["LM_BENCH_E2E_SPLIT"]) / 256.0)) * 256
        sizes = [first, n - first]
    stepper = ShardedHostStepper(pipe.cfg, n, pipe.policy, shards=shards, device=device, seed=1000 * rank + 42,
                                 env_id_offset=rank * n, norm_mean=pipe.norm_mean if mlp else None,
                                 norm_inv_std=pipe.norm_inv_std if mlp else None, readback_obs=readback_obs, shard_sizes=sizes)
    sets = []                                     # STATE_SETS x shards pinned host states (> L2, and never the same twice in a row)
    for i in range(STATE_SETS):
        sets.append([s.scene.pack_host(s.scene.generate(i)) for s in stepper.shards])
    if os.environ.get("LM_BENCH_WC"):             # experiment: write-combined pinned memory for the state on its way to the device
        from locomanipulationrl_b200._lib import HostBuffer
        for per_shard in sets:
            for hs in per_shard:
                for grp in list(hs):
                    hb = HostBuffer(hs[grp].numel(), write_combined=True)
                    hb.tensor.copy_(hs[grp].view(-1))
                    hs[grp] = hb
    g = torch.Generator().manual_seed(11 + rank)
    for s in stepper.shards:
        s.uniforms_host.copy_(torch.rand(s.n, 3, generator=g))
    h2d = stepper.h2d_bytes + sum(t.numel() * 4 for hs in sets[0] for t in hs.values())
    d2h = stepper.d2h_bytes
    steps = max(4, min(args.steps, 60))
    for t in range(3):
        stepper.step(sets[t % STATE_SETS])
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    t0 = time.perf_counter()
    for t in range(steps):
        stepper.step(sets[t % STATE_SETS])
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(dt, op=tdist.ReduceOp.MAX)
    return {"value": n * world * steps / float(dt.item()), "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "ms_per_step": float(dt.item()) / steps * 1e3,
            "api": f"ShardedHostStepper.step: {shards} env shard(s) x lm_task_step_host_policy_async (env step + policy forward + read-back in one "
                   "C-ABI call per shard; the shards' physics-state copies queue on one copy stream), pinned host buffers, "
                   "host sync every step; H2D actions, goal uniforms, physics state; D2H next actions, value, reward, reset flags"
                   + (", observations" if readback_obs else " (observations stay on the device for the device-resident policy)")}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The contract is ONE JSON line on stdout: libraries (NCCL's version banner) also write there, so everything but this
    line is diverted to stderr (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                              # fd 1 -> stderr for everything except emit()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cotrain64k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch eagerly (for ncu captures)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--comm-stream", default="main", choices=["main", "side"],
                    help="where the per-horizon stat + gradient all-reduce runs (N > 1)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="rollout-boundary all-reduce: our one-shot NVLink peer-memory kernel (default) or torch.distributed/NCCL")
    ap.add_argument("--no-train", action="store_true", help="skip the train block (rollout + GAE + PPO update)")
    ap.add_argument("--e2e-shards", type=int, default=2,
                    help="env shards the host-buffer step is pipelined over: the shards' state copies queue on one copy stream, shard k's "
                         "kernels and read-back run under shard k+1's copy")
    ap.add_argument("--no-large", action="store_true", help="skip the 524,288-env k_post roofline line")
    ap.add_argument("--no-kernel-times", action="store_true",
                    help="skip the per-kernel timing section (ncu launch lists: the capture then holds the warm-up and timed steps only)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
