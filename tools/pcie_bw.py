"""Pinned host <-> device copy bandwidth of the box (the bound of bench.py's e2e leg)."""
import time, torch
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
h2d = t(lambda: d.copy_(h, non_blocking=True)); d2h = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
bi = t(both)
print(f"H2D {n/h2d/1e9:.1f} GB/s  D2H {n/d2h/1e9:.1f} GB/s  both directions at once {2*n/bi/1e9:.1f} GB/s total")
