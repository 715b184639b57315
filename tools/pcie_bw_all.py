"""Concurrent pinned host <-> device copy bandwidth on every GPU of the box (the bound of the N-GPU e2e leg).
   python -m torch.distributed.run --nproc-per-node N tools/pcie_bw_all.py"""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=4):
    fn(); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps; dist.barrier(); return dt
h2d = t(lambda: d.copy_(h, non_blocking=True)); d2h = t(lambda: h2.copy_(d2, non_blocking=True))
res = torch.tensor([n / h2d / 1e9, n / d2h / 1e9]); allr = [torch.zeros(2) for _ in range(dist.get_world_size())]
dist.all_gather(allr, res)
if dist.get_rank() == 0:
    print("per GPU H2D GB/s:", [round(float(r[0]), 1) for r in allr], "sum", round(sum(float(r[0]) for r in allr), 1))
    print("per GPU D2H GB/s:", [round(float(r[1]), 1) for r in allr], "sum", round(sum(float(r[1]) for r in allr), 1))
    print("cores", os.cpu_count())
