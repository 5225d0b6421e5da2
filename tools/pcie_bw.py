import torch, time
n = 1<<30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
for _ in range(2): d.copy_(h, non_blocking=True); h2.copy_(d2, non_blocking=True)
print("H2D GB/s", n/1e9/t(lambda: d.copy_(h, non_blocking=True)))
print("D2H GB/s", n/1e9/t(lambda: h2.copy_(d2, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidir each GB/s", n/1e9/t(both))
