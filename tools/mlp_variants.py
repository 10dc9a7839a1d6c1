"""Times lm_mlp_forward at 65,536 envs for every liblm_b200*.so variant given on the command line (one subprocess per library so that
each gets a fresh load): python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/*.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from locomanipulationrl_b200.policy.mlp import MlpActorCritic
torch.manual_seed(0)
n = 65536
m = MlpActorCritic(device="cuda:0")
sets = [torch.randn(n, 64, device="cuda").clamp(-5, 5) for _ in range(8)]          # 8 x 16.8 MB > L2
mu, v = torch.empty(n, 12, device="cuda"), torch.empty(n, 1, device="cuda")
for s in sets: m(s, out=(mu, v))
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for s in sets: m(s, out=(mu, v))
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(25): g.replay()
b.record(); torch.cuda.synchronize()
with torch.no_grad():
    ref = m.reference_forward(sets[-1])[0]
print(f"{a.elapsed_time(b) / (25 * 8) * 1e3:.2f} us/launch   max |mu - fp32| {float((mu - ref).abs().max()):.2e}")
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, LM_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), (out.stdout.strip() or out.stderr.strip()[-400:]), flush=True)
