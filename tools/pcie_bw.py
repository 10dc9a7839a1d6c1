"""Diagnostic: pinned host<->device copy bandwidth on this box (alone and both directions at once)."""
import time
import torch
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 2)
print(f"H2D alone {run(True, False):.1f} GB/s; D2H alone {run(False, True):.1f} GB/s; both at once {run(True, True):.1f} GB/s each")
for sz in (1 << 20, 4 << 20, 16 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        d1[:sz].copy_(h1[:sz], non_blocking=True)
    torch.cuda.synchronize(); print(f"H2D {sz >> 20} MiB chunks: {sz * 50 / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
