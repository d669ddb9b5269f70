import torch, time
x = torch.empty(20 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(20 << 20, dtype=torch.uint8, device="cuda")
for n in (1, 3):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        d.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 20
    print("H2D 20MB pinned ms", dt * 1e3, "GB/s", (20 << 20) / dt * 1e-9)
y = torch.empty(3 << 20, dtype=torch.uint8).pin_memory()
dd = torch.empty(3 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20):
    y.copy_(dd, non_blocking=True); torch.cuda.synchronize()
print("D2H 3MB pinned ms", (time.perf_counter() - t) / 20 * 1e3)
z = torch.empty(3 << 20, dtype=torch.uint8)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20):
    z.copy_(dd); torch.cuda.synchronize()
print("D2H 3MB pageable ms", (time.perf_counter() - t) / 20 * 1e3)
