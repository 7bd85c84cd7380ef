import torch, time
x = torch.empty(84_688_896 // 4, dtype=torch.float32).pin_memory()
y = torch.empty(56_459_264 // 4, dtype=torch.float32).pin_memory()
xd = torch.empty_like(x, device="cuda"); yd = torch.empty_like(y, device="cuda")
for _ in range(3):
    xd.copy_(x, non_blocking=True); y.copy_(yd, non_blocking=True)
torch.cuda.synchronize()
def t(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
a = t(lambda: xd.copy_(x, non_blocking=True)); b = t(lambda: y.copy_(yd, non_blocking=True))
print(f"H2D 84.7 MB: {a*1e3:.3f} ms = {84.69/a/1e3:.1f} GB/s; D2H 56.5 MB: {b*1e3:.3f} ms = {56.46/b/1e3:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): xd.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(yd, non_blocking=True)
c = t(both)
print(f"both directions at once: {c*1e3:.3f} ms")
