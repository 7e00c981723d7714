"""ORACLE — test infrastructure only.  CPU restatement of exact k-NN search.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this package; the product path (``cuvs-rag_b200/``) never does.

Parity status: **unpinned at the ANN boundary** — the reference repo holds no golden distances,
ids or recall for its cuVS / FAISS calls; every reference test mocks them.  The
arithmetic lives in un-vendored third-party wheels (``cuvs-cu12==25.6.0``,
``Attempt_1/pip-requirements.txt:9``; ``faiss-gpu`` 1.7.2, ``Latest/faiss.ipynb`` cell 0), so
this file restates their *published* algorithm:

* ``exact_knn`` follows FAISS ``IndexFlatL2`` / ``IndexFlatIP`` for batches >= 20 queries
  (``Latest/faiss-main.ipynb`` cells 9-10 is the reference call site): blocked fp32 sgemm,
  ``||q||^2 + ||x||^2 - 2 q.x`` for L2 (clamped at 0), a per-query heap (here ``torch.topk``),
  results best-first (ascending L2, descending IP).
* ``sklearn_brute_knn`` is the reference's own CPU baseline, runnable here
  (``Attempt_1/VectorSearch_QuestionRetrieval.ipynb:L878``:
  ``NearestNeighbors(algorithm='brute', n_jobs=-1)``); it pins ``exact_knn`` in the tests.
* ``pairwise_f64`` is the float64 ground truth used to adjudicate ties (north star: ids equal
  except for distance ties within 1e-3 relative).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def round_through(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 copy of ``x`` holding exactly the values a ``dtype`` tensor holds."""
    return x.to(dtype).to(torch.float32)


def exact_knn(db: torch.Tensor, queries: torch.Tensor, k: int, metric: str = "sqeuclidean",
              block: int = 65536) -> Tuple[torch.Tensor, torch.Tensor]:
    """Blocked fp32 exact search.  db [N,D], queries [Q,D] (any float dtype; promoted to fp32).

    Returns (distances [Q,k] fp32, ids [Q,k] int64) best-first.  Rows are processed in blocks of
    ``block`` with a running top-k, as FAISS does; missing results (k > N) are (inf|-inf, -1).
    """
    db = db.to(torch.float32)
    q = queries.to(torch.float32)
    n, nq = db.shape[0], q.shape[0]
    l2 = metric in ("sqeuclidean", "l2", "L2", "euclidean")
    best_d = torch.full((nq, k), float("inf"))
    best_i = torch.full((nq, k), -1, dtype=torch.int64)
    qn = (q * q).sum(1, keepdim=True) if l2 else None
    for s in range(0, n, block):
        xb = db[s:s + block]
        ip = q @ xb.T
        if l2:
            score = (qn + (xb * xb).sum(1)[None, :] - 2.0 * ip).clamp_(min=0.0)
        else:
            score = -ip
        cat_d = torch.cat([best_d, score], 1)
        cat_i = torch.cat([best_i, torch.arange(s, s + xb.shape[0])[None, :].expand(nq, -1)], 1)
        kk = min(k, cat_d.shape[1])
        d, pos = torch.topk(cat_d, kk, dim=1, largest=False, sorted=True)
        best_d[:, :kk] = d
        best_i[:, :kk] = torch.gather(cat_i, 1, pos)
    if not l2:
        best_d = -best_d
        best_d[best_i < 0] = float("-inf")
    return best_d, best_i


def pairwise_f64(db: torch.Tensor, queries: torch.Tensor, metric: str = "sqeuclidean") -> torch.Tensor:
    """Full [Q,N] float64 distance (L2^2) or similarity (IP) matrix — small inputs only."""
    x = db.to(torch.float64)
    q = queries.to(torch.float64)
    if metric in ("sqeuclidean", "l2", "L2", "euclidean"):
        return torch.cdist(q, x, p=2.0) ** 2
    return q @ x.T


def sklearn_brute_knn(db: np.ndarray, queries: np.ndarray, k: int, metric: str = "sqeuclidean"):
    """The reference's CPU path (VectorSearch_QuestionRetrieval.ipynb:L878), euclidean variant."""
    from sklearn.neighbors import NearestNeighbors

    if metric not in ("sqeuclidean", "l2", "L2", "euclidean"):
        raise ValueError("sklearn brute baseline is restated for the euclidean metric only")
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean", n_jobs=-1)
    nn.fit(db)
    d, i = nn.kneighbors(queries)
    return (d ** 2).astype(np.float32), i.astype(np.int64)


def topk_parity_report(got_d: torch.Tensor, got_i: torch.Tensor, db: torch.Tensor,
                       queries: torch.Tensor, k: int, metric: str = "sqeuclidean",
                       rtol: float = 1e-3) -> dict:
    """Check a [Q,k] result against float64 ground truth with the north-star tie rule.

    A result row passes when (a) ids are unique and in range, (b) every id that is strictly better
    than the k-th true distance by more than the tie window is present, (c) every returned id's
    true distance is within the tie window of the k-th true distance, (d) reported distances match
    the true distances of the returned ids to ``rtol`` (relative to the distance scale).
    """
    l2 = metric in ("sqeuclidean", "l2", "L2", "euclidean")
    full = pairwise_f64(db, queries, metric)
    score = full if l2 else -full
    n = db.shape[0]
    kk = min(k, n)
    ref_sorted, ref_idx = torch.sort(score, dim=1, stable=True)
    kth = ref_sorted[:, kk - 1]
    scale = score.abs().mean(dim=1).clamp_min(1e-12)
    window = rtol * torch.maximum(kth.abs(), scale)
    gi = got_i[:, :kk].to(torch.int64).cpu()
    gd = got_d[:, :kk].to(torch.float64).cpu()
    bad_range = int(((gi < 0) | (gi >= n)).sum())
    gi_c = gi.clamp(0, n - 1)
    got_true = torch.gather(score, 1, gi_c)
    dup = sum(int(len(set(r.tolist())) != kk) for r in gi)
    too_far = int((got_true > (kth + window)[:, None]).sum())
    must = score < (kth - window)[:, None]
    present = torch.zeros_like(must)
    present.scatter_(1, gi_c, True)
    missed = int((must & ~present).sum())
    gd_score = gd if l2 else -gd
    derr = ((gd_score - got_true).abs() / torch.maximum(got_true.abs(), scale[:, None])).max().item()
    sorted_ok = bool((gd_score[:, 1:] >= gd_score[:, :-1] - 1e-6 * scale[:, None]).all())
    exact_match = float((gi == ref_idx[:, :kk]).float().mean())
    return {"bad_range": bad_range, "duplicates": dup, "too_far": too_far, "missed": missed,
            "max_rel_dist_err": derr, "sorted": sorted_ok, "exact_id_match": exact_match,
            "ok": bad_range == 0 and dup == 0 and too_far == 0 and missed == 0 and derr <= rtol
            and sorted_ok}
