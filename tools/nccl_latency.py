"""Diagnostic (run under torchrun): latency of the two per-horizon all-reduces (58,637 fp32 gradients, 14 fp64 stats)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from locomanipulationrl_b200 import dist as lmdist
rank, world, local = lmdist.init_from_env("nccl")
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
g = torch.zeros(58637, device=dev); s = torch.zeros(14, dtype=torch.float64, device=dev)
def both():
    dist.all_reduce(s); dist.all_reduce(g); g.div_(world)
for _ in range(5): both()
torch.cuda.synchronize(); dist.barrier()
def timeit(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_e = timeit(both)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): both()
t_g = timeit(gr.replay)
t_s = timeit(lambda: dist.all_reduce(s)); t_gr = timeit(lambda: dist.all_reduce(g))
if rank == 0: print(f"world {world}: eager pair {t_e:.1f} us, graph pair {t_g:.1f} us, stats only {t_s:.1f} us, grads only {t_gr:.1f} us", file=sys.stderr)
dist.destroy_process_group()
