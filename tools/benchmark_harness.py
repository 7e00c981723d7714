"""Benchmark harness in the shape the reference's users run:
multi-k recall ([1,5,10,50,100,500,1000,2000], improved_multi_gpu_rag.py:37-48), top-2K retrieval,
batch-size sweep and sharded-vs-replicated comparison (Latest/faiss.ipynb cells 10-11), results as
CSV (colab_a100_test.ipynb cell 23).  Everything goes through the second-generation drop-in API
(ParallelIndexBuilder / ParallelSearchEngine / RecallEvaluator).

  python tools/benchmark_harness.py [--n 1000000] [--dim 768] [--index ivf_flat|ivf_pq|faiss_flat]
                                    [--gpus G] [--out gpurun_out/harness.csv]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

import cuvs_rag_b200 as b2
from cuvs_rag_b200.improved_multi_gpu_rag import (IndexType, ParallelIndexBuilder, ParallelSearchEngine,
                                                  RecallEvaluator, SearchConfig)


def clustered(n, d, n_comp, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    cent = torch.randn(n_comp, d, generator=g, device=dev)
    out = torch.empty((n, d), dtype=torch.float16, device=dev)
    for s in range(0, n, 1 << 19):
        e = min(n, s + (1 << 19))
        lab = torch.randint(0, n_comp, (e - s,), generator=g, device=dev)
        out[s:e] = (cent[lab] + 0.42 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.float16)
    return out


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--index", default="ivf_flat", choices=["ivf_flat", "ivf_pq", "faiss_flat"])
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--n-lists", type=int, default=1024)
    ap.add_argument("--n-probes", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/harness.csv")
    args = ap.parse_args()
    G = max(1, min(args.gpus, torch.cuda.device_count()))
    itype = IndexType(args.index)
    params = {"n_lists": args.n_lists, "pq_dim": max(16, args.dim // 8), "kmeans_n_iters": 10}
    rows = []

    # ---- data: one corpus, row-sharded over G GPUs (and a full copy per GPU for "replicated")
    full = clustered(args.n, args.dim, args.n_lists, torch.device("cuda:0"), 7)
    bounds = [args.n * i // G for i in range(G + 1)]
    parts = [full[bounds[i]:bounds[i + 1]].to(f"cuda:{i}") for i in range(G)]
    nq_max = 10_000
    gq = torch.Generator(device="cuda:0").manual_seed(11)
    qi = torch.randint(0, args.n, (nq_max,), generator=gq, device="cuda:0")
    queries = (full[qi].float() + 0.1 * torch.randn((nq_max, args.dim), generator=gq, device="cuda:0")).to(torch.float16)

    builder = ParallelIndexBuilder(G)
    res = builder.build_indices_parallel(parts, itype, params)
    assert res["success"], res
    rows.append({"section": "build", "index": args.index, "gpus": G, "n": args.n, "dim": args.dim,
                 "avg_build_s": res["avg_time"], "total_build_s": res["total_time"]})
    search_params = {"n_probes": args.n_probes, "refine_ratio": 4 if itype == IndexType.IVF_PQ else 0}

    # ---- exact ground truth (top-2000) from a flat index over the same shards
    truth_res = builder.build_indices_parallel(parts, IndexType.FAISS_FLAT, {})
    truth_eng = ParallelSearchEngine(truth_res["indexes"], IndexType.FAISS_FLAT, SearchConfig(top_k=2000))
    n_eval = 200
    _, truth_ids = truth_eng.parallel_search(queries[:n_eval])

    # ---- multi-k recall of the top-2K retrieval (IVF-PQ serves k <= 128: evaluated at 100)
    top_k = 2000 if itype != IndexType.IVF_PQ else 100
    cfg = SearchConfig(top_k=top_k)
    eng = ParallelSearchEngine(res["indexes"], itype, cfg)
    d, ids = eng._search_merged(queries[:n_eval], top_k, search_params)
    ids = ids.cpu().numpy()
    for k in [k for k in cfg.recall_k_values if k <= top_k]:
        rec = np.mean([RecallEvaluator.calculate_recall_at_k(ids[j], truth_ids[j][:k], k) for j in range(n_eval)])
        rows.append({"section": "recall", "index": args.index, "gpus": G, "k": k, "top_k": top_k,
                     "n_probes": args.n_probes, "recall": float(rec)})

    # ---- batch-size sweep at k = 10 (sharded)
    for bs in [1, 10, 100, 1000, 10_000]:
        sec, _ = timed(lambda: eng._search_merged(queries[:bs], 10, search_params))
        rows.append({"section": "batch_sweep", "index": args.index, "gpus": G, "mode": "sharded",
                     "batch": bs, "k": 10, "ms": sec * 1e3, "qps": bs / sec})

    # ---- sharded vs replicated (faiss.ipynb cell 11: co.shard = True / False)
    if G > 1:
        rep_parts = [full.to(f"cuda:{i}") for i in range(G)]
        rep = [b2.NativeIndex.flat(p) if itype == IndexType.FAISS_FLAT else
               (b2.NativeIndex.ivf_flat(p, args.n_lists, kmeans_iters=10) if itype == IndexType.IVF_FLAT else
                b2.NativeIndex.ivf_pq(p, args.n_lists, params["pq_dim"], kmeans_iters=10)) for p in rep_parts]

        def replicated(bs):
            # every GPU holds the whole index and answers a slice of the batch
            outs = []
            for i in range(G):
                lo, hi = bs * i // G, bs * (i + 1) // G
                if hi > lo:
                    outs.append(rep[i].search(queries[lo:hi].to(f"cuda:{i}", non_blocking=True), 10,
                                              n_probes=args.n_probes, refine_ratio=search_params["refine_ratio"]))
            for i in range(G):
                torch.cuda.synchronize(i)
            return outs

        for bs in [1, 100, 10_000]:
            sec, _ = timed(lambda: replicated(bs))
            rows.append({"section": "batch_sweep", "index": args.index, "gpus": G, "mode": "replicated",
                         "batch": bs, "k": 10, "ms": sec * 1e3, "qps": bs / sec})

    df = pd.DataFrame(rows)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    df.to_csv(args.out, index=False)
    print(df.to_string(index=False))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
