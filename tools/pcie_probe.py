import torch, time
B, n, ld = 4096, 613, 616
xd = torch.zeros((B, ld), dtype=torch.float64, device="cuda")
xh2 = torch.empty((B, n), dtype=torch.float64).pin_memory()
xh1 = torch.empty((B, ld), dtype=torch.float64).pin_memory()
s = torch.cuda.Stream()
def t(fn, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(reps): fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for _ in range(2):
    a = t(lambda: xh2.copy_(xd[:, :n], non_blocking=True))
    b = t(lambda: xh1.copy_(xd, non_blocking=True))
    print("2D D2H %.3f ms (%.1f GB/s) | 1D D2H %.3f ms (%.1f GB/s)" % (a, B*n*8/a/1e6, b, B*ld*8/b/1e6))
zc = torch.empty((B, 100, 2), dtype=torch.float64).pin_memory(); zd = torch.empty_like(zc, device="cuda")
c = t(lambda: zd.copy_(zc, non_blocking=True))
print("H2D 6.5 MB %.3f ms (%.1f GB/s)" % (c, zc.numel()*8/c/1e6))
