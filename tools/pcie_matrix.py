"""Host <-> device copy bandwidth of the box for every GPU count the bench runs at, ONE process driving the GPUs
(the bound of the N-GPU e2e legs; profiles/r2_pcie.md is made from its output).
    python tools/pcie_matrix.py            # N = 1, 2, 4, 8 (as many as the box has), plus pairs that tell a shared uplink
Prints one JSON object: per device set, the aggregate H2D / D2H / both-directions GB/s with all of the set copying at once."""
import json, subprocess, time
import torch

n_dev = torch.cuda.device_count()
SZ = 1 << 30
bufs = {}
for d in range(n_dev):
    torch.cuda.set_device(d)
    bufs[d] = dict(h=torch.empty(SZ, dtype=torch.uint8, pin_memory=True).fill_(1), h2=torch.empty(SZ, dtype=torch.uint8, pin_memory=True),
                   d=torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{d}"), d2=torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{d}"),
                   s1=torch.cuda.Stream(device=d), s2=torch.cuda.Stream(device=d))


def run(devs, h2d, d2h, reps=3):
    def once():
        for d in devs:
            b = bufs[d]
            if h2d:
                with torch.cuda.stream(b["s1"]):
                    b["d"].copy_(b["h"], non_blocking=True)
            if d2h:
                with torch.cuda.stream(b["s2"]):
                    b["h2"].copy_(b["d2"], non_blocking=True)
    def sync():
        for d in devs:
            torch.cuda.synchronize(d)
    once(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    sync()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * SZ * (int(h2d) + int(d2h)) / dt / 1e9


out = {"gpus": n_dev, "sets": {}}
sets = [list(range(k)) for k in (1, 2, 4, 8) if k <= n_dev]
if n_dev >= 8:
    sets += [[0, 4], [0, 2], [0, 1, 4, 5], [0, 2, 4, 6]]
for devs in sets:
    out["sets"][",".join(map(str, devs))] = {"h2d_GBps": round(run(devs, True, False), 1), "d2h_GBps": round(run(devs, False, True), 1),
                                             "both_GBps": round(run(devs, True, True), 1)}
try:
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    out["numa"] = subprocess.run(["bash", "-c", "lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'; free -g | head -2"], capture_output=True, text=True, timeout=30).stdout
except Exception as e:
    out["topo"] = str(e)
print(json.dumps(out))
