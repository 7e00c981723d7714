import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

def clustered(n, d, c, seed, sigma=0.2):
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(c, d, generator=g)
    return cent[torch.randint(0, c, (n,), generator=g)] + sigma * torch.randn(n, d, generator=g)

for (n, d, ncl, dt) in [(50000, 64, 32, torch.float16), (100000, 128, 256, torch.float16),
                        (50000, 64, 32, torch.float32), (50000, 64, 32, torch.bfloat16),
                        (60000, 768, 64, torch.float16)]:
    x = clustered(n, d, 32, 1).to(dt).cuda()
    for iters in (0, 1, 2, 10):
        c, lab = b2.kmeans_fit(x, ncl, iters=iters, seed=3)
        torch.cuda.synchronize()
        xf = x.float()
        # reference assignment against the RETURNED centroids (rounded as the engine rounds them)
        cr = c if dt == torch.float32 else c.to(dt).float()
        dist = (cr * cr).sum(1)[None, :] - 2.0 * xf @ cr.T
        ref = dist.argmin(1)
        agree = (ref == lab.long()).float().mean().item()
        # how much worse is the chosen centroid than the best one
        chosen = dist.gather(1, lab.long().clamp(0, ncl - 1)[:, None])[:, 0]
        gap = (chosen - dist.min(1).values).max().item()
        cnt = torch.bincount(lab.long().clamp(0, ncl - 1), minlength=ncl)
        inertia = ((xf - c[lab.long()]) ** 2).sum(1).mean().item()
        print(json.dumps({"n": n, "d": d, "ncl": ncl, "dt": str(dt), "iters": iters, "agree": round(agree, 5),
                          "max_gap": gap, "min_lab": int(lab.min()), "max_lab": int(lab.max()),
                          "max_cnt": int(cnt.max()), "min_cnt": int(cnt.min()), "inertia": round(inertia, 3),
                          "cent_nan": bool(torch.isnan(c).any()), "cent_absmax": c.abs().max().item()}))
