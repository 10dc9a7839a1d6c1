"""Diagnostic: where one host-buffer step (ShardedHostStepper.step) spends its time.  CUDA events on every shard stream and on the
copy stream, relative to the step's start; host enqueue time; a pure-H2D reference of the same buffers.  Needs a B200.
    python tools/e2e_probe.py [shards ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import configs  # noqa: E402
from locomanipulationrl_b200.host_step import ShardedHostStepper  # noqa: E402
from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402

n = 65536
cfg = configs.get_task_config("JointLocomanipulationVertical")
dev = torch.device("cuda:0")
torch.manual_seed(0)
policy = MlpActorCritic(device=dev)
policy.pack()
mean, inv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
for shards in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    for copy_stream in ([False] if shards == 1 else [True, False]):
        st = ShardedHostStepper(cfg, n, policy, shards=shards, device=dev, norm_mean=mean, norm_inv_std=inv, readback_obs=False,
                                copy_stream=copy_stream)
        sets = [[s.scene.pack_host(s.scene.generate(i)) for s in st.shards] for i in range(6)]
        for t in range(4):
            st.step(sets[t % 6])
        torch.cuda.synchronize()
        # wall clock
        t0 = time.perf_counter()
        for t in range(40):
            st.step(sets[t % 6])
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 40
        # one instrumented step: events before / after each shard's call on its stream (and on the copy stream)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        start = ev()
        start.record(torch.cuda.current_stream())
        for s in st.shards:
            s.stream.wait_event(start)
        if st.h2d_stream is not None:
            st.h2d_stream.wait_event(start)
        ends, copies = [], []
        real_sync = st.lib.lm_stream_synchronize
        h0 = time.perf_counter()
        p = st.parity
        import ctypes as C
        from locomanipulationrl_b200 import _lib
        for s, hs in zip(st.shards, sets[0]):
            io, task = s.io, s.task
            io.actions_host = s.mu_host[p].data_ptr()
            for grp in ("loco", "mani"):
                if grp in hs:
                    setattr(io, grp + "_state_host", hs[grp].data_ptr()); setattr(io, grp + "_state_floats", hs[grp].numel())
                    setattr(io, grp + "_state_dev", s.scene.arena[grp].data_ptr())
            s.pol.mu_host = s.mu_host[p ^ 1].data_ptr()
            _lib.check(st.lib.lm_task_step_host_policy_async(task._handle, task._lp(), task._mp(), task._env_pos.data_ptr(), C.byref(task._buffers),
                                                             C.byref(io), C.byref(s.pol), s.stream.cuda_stream), "step")
            e = ev(); e.record(s.stream); ends.append(e)
            if st.h2d_stream is not None:
                c = ev(); c.record(st.h2d_stream); copies.append(c)
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        st.parity = p ^ 1
        print(f"shards {shards} copy_stream {copy_stream}: wall {wall * 1e3:.3f} ms/step; host enqueue {(h1 - h0) * 1e3:.3f} ms; "
              f"shard done at {[round(start.elapsed_time(e), 3) for e in ends]} ms; state landed at {[round(start.elapsed_time(c), 3) for c in copies]} ms",
              flush=True)
        del st
# pure H2D of one step's inputs, one stream, same pinned buffers
st = ShardedHostStepper(cfg, n, policy, shards=1, device=dev, norm_mean=mean, norm_inv_std=inv, readback_obs=False)
hs = st.shards[0].scene.pack_host(st.shards[0].scene.generate(0))
s = st.shards[0]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    for grp in hs:
        s.scene.arena[grp][: hs[grp].numel()].copy_(hs[grp], non_blocking=True)
    s.actions_dev.copy_(s.mu_host[0], non_blocking=True)
b.record(); torch.cuda.synchronize()
nbytes = sum(hs[g].numel() * 4 for g in hs) + s.actions_dev.numel() * 4
print(f"pure H2D of {nbytes / 1e6:.1f} MB: {a.elapsed_time(b) / 20:.3f} ms ({nbytes / (a.elapsed_time(b) / 20 * 1e-3) / 1e9:.1f} GB/s)")
