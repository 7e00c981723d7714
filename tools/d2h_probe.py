import time, torch, numpy as np
d = torch.rand(10000, 100, device="cuda"); i = torch.randint(0, 10**7, (10000, 100), device="cuda")
torch.cuda.synchronize()
def t(f, n=20):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("cpu().numpy()      ms", t(lambda: (d.cpu().numpy(), i.cpu().numpy())))
pd = torch.empty(d.shape, dtype=d.dtype, pin_memory=True); pi = torch.empty(i.shape, dtype=i.dtype, pin_memory=True)
def pinned():
    pd.copy_(d, non_blocking=True); pi.copy_(i, non_blocking=True); torch.cuda.current_stream().synchronize()
    return np.array(pd.numpy()), np.array(pi.numpy())
print("pinned + np.copy   ms", t(pinned))
def pinned_view():
    pd.copy_(d, non_blocking=True); pi.copy_(i, non_blocking=True); torch.cuda.current_stream().synchronize()
    return pd.numpy(), pi.numpy()
print("pinned view        ms", t(pinned_view))
def pinned_fresh():
    a = torch.empty(d.shape, dtype=d.dtype, pin_memory=True); b = torch.empty(i.shape, dtype=i.dtype, pin_memory=True)
    a.copy_(d, non_blocking=True); b.copy_(i, non_blocking=True); torch.cuda.current_stream().synchronize()
    return a.numpy(), b.numpy()
print("fresh pinned (cached alloc) ms", t(pinned_fresh))
q = torch.randn(10000, 768).to(torch.bfloat16).pin_memory()
print("H2D pinned 15MB    ms", t(lambda: q.to("cuda", non_blocking=True)))
