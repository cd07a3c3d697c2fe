import torch, time
a = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("H2D", lambda: d.copy_(a, non_blocking=True)), ("D2H", lambda: a.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    print(name, 128 * 10 / 1024 / (time.perf_counter() - t), "GiB/s")
b = torch.empty(128 << 20, dtype=torch.uint8).pin_memory(); e = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): b.copy_(e, non_blocking=True)
torch.cuda.synchronize()
print("bidirectional", 256 * 10 / 1024 / (time.perf_counter() - t), "GiB/s total")
