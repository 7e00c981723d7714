"""ORACLE — test infrastructure only.  CPU restatement of IVF-Flat / IVF-PQ semantics.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.

Parity status: **unpinned at the cuVS boundary, pinned on scikit-learn** — the reference holds no
golden ids / recall for its ``cuvs.neighbors.ivf_flat`` / ``ivf_pq`` calls
(``index_building_coordinator.py:392-404``, ``improved_multi_gpu_rag.py:126-138, 225-233``: every
reference test mocks them); cuVS 25.06 itself is an un-vendored wheel
(``Attempt_1/pip-requirements.txt:9``).  This file restates the published algorithm those calls
implement.  Each step of the restatement is anchored on scikit-learn - the library behind the
reference's own CPU baseline (``VectorSearch_QuestionRetrieval.ipynb:L878``) - through
``tests/golden/ivf.npz`` (generator: ``tests/golden/make_golden_ivf.py``): Lloyd iterations from a
fixed init == ``sklearn.cluster.KMeans``, list assignment == ``KMeans.predict``, probe selection ==
``NearestNeighbors`` over the centroids, the list scan == ``NearestNeighbors`` over the probed rows,
the residual-PQ ADC distance == the sum of per-subspace squared distances (``tests/test_oracle.py``).
The GPU path is compared with this oracle through recall at identical ``n_lists`` / ``n_probes``
(the north-star criterion), not bit-for-bit (k-means is seeded differently and sums in a
different order):

* coarse quantizer: Lloyd k-means (``kmeans_n_iters`` iterations, default 20) on a strided
  subsample (``kmeans_trainset_fraction``, default 0.5), then every row joins its nearest
  centroid's list;
* IVF-Flat search: the ``n_probes`` best centroids per query (default 20), exact distances to
  every row of those lists, best k;
* IVF-PQ: residual to the list centroid, split in ``pq_dim`` sub-vectors, 256-entry codebook per
  sub-space (k-means on residual sub-vectors), asymmetric distance via a per-(query, list) table.
"""
from __future__ import annotations

from typing import Tuple

import torch


ADJUST_WEIGHT = 7.0       # weight of the large cluster's centre against the picked data row


def balance_pairs(counts, n):
    """WHICH clusters move: clusters above 1.5x the average size want floor(size/avg) - 1 extra
    centroids, taken from the smallest clusters below 0.5x the average (smallest first, largest
    cluster served first).  Returns donor_of[c] (-1 = keep).  cuVS's adjust_centers picks the large
    cluster through a random data row (probability ~ size) and only moves clusters below a quarter
    of the average; ranking instead empties over-full "hub" lists first (measured: DESIGN.md §4)."""
    ncl = len(counts)
    donor = [-1] * ncl
    avg = n / float(ncl)
    order = sorted(range(ncl), key=lambda c: (counts[c], c))
    lo, hi = 0, ncl - 1
    while lo < hi:
        big = order[hi]
        if counts[big] <= 1.5 * avg:
            break
        quota = max(1, int(counts[big] / avg) - 1)
        while quota > 0 and lo < hi and counts[order[lo]] < 0.5 * avg:
            donor[order[lo]] = big
            lo += 1
            quota -= 1
        if quota > 0:
            break
        hi -= 1
    return donor


def adjust_centers(cent: torch.Tensor, cnt: torch.Tensor, lab: torch.Tensor, x: torch.Tensor,
                   g: torch.Generator) -> torch.Tensor:
    """WHERE they move: the re-seed position of cuVS / RAFT balanced k-means ("adjust_centers",
    published in raft/cluster/detail/kmeans_balanced.cuh) restated - next to the large cluster l,
    nudged towards one of its members i:

        centre[small] = (wc * centre[l] + x[i]) / (wc + 1),   wc = min(size[l], ADJUST_WEIGHT)

    so the next assignment splits the large cluster.  (Restarting ON a data row - this file's
    round-1 reading - isolates the new centre in high dimensions: ||x||^2 dominates its score and
    only the row itself joins, leaving singleton lists: 60 % of the lists on iid Gaussian rows.)"""
    n = x.shape[0]
    out = cent.clone()
    donor = balance_pairs(cnt.tolist(), n)
    for c, l in enumerate(donor):
        if l < 0 or l == c:
            continue
        rows = torch.nonzero(lab == l)[:, 0]
        i = rows[int(torch.randint(0, rows.numel(), (1,), generator=g))]
        wc = min(float(cnt[l]), ADJUST_WEIGHT)
        out[c] = (wc * cent[l] + x[i]) / (wc + 1.0)
    return out


def kmeans(x: torch.Tensor, n_clusters: int, iters: int = 20, seed: int = 0,
           balance: bool = True, init: "torch.Tensor | None" = None) -> torch.Tensor:
    """Lloyd iterations (assign, mean) with a balancing step between them: after the update of
    every iteration but the last, under-full clusters are re-seeded inside over-full ones
    (``balance_pairs`` + ``adjust_centers``; cuVS adjusts before the E-step of every iteration
    after the first - the same schedule).  A cluster that is empty when no balancing follows keeps
    its previous centre."""
    x = x.to(torch.float32)
    n = x.shape[0]
    g = torch.Generator().manual_seed(seed)
    cent = x[torch.randperm(n, generator=g)[:n_clusters]].clone() if init is None \
        else init.to(torch.float32).clone()
    for it in range(iters):
        lab = assign(x, cent)
        sums = torch.zeros_like(cent).index_add_(0, lab, x)
        cnt = torch.bincount(lab, minlength=n_clusters)
        mean = sums / cnt.clamp_min(1).to(torch.float32)[:, None]
        cent = torch.where((cnt > 0)[:, None], mean, cent)
        if balance and it + 1 < iters and n_clusters > 1:
            cent = adjust_centers(cent, cnt, lab, x, g)
    return cent


def assign(x: torch.Tensor, cent: torch.Tensor, block: int = 65536) -> torch.Tensor:
    cn = (cent * cent).sum(1)
    out = torch.empty(x.shape[0], dtype=torch.int64)
    for s in range(0, x.shape[0], block):
        xb = x[s:s + block].to(torch.float32)
        out[s:s + block] = (cn[None, :] - 2.0 * (xb @ cent.T)).argmin(1)
    return out


class IvfFlatOracle:
    def __init__(self, db: torch.Tensor, n_lists: int, metric: str = "sqeuclidean",
                 iters: int = 20, train_fraction: float = 0.5, seed: int = 0, balance: bool = True,
                 init=None):
        self.metric = metric
        self.db = db.to(torch.float32)
        stride = max(1, int(1.0 / train_fraction + 1e-6))
        train = self.db[::stride]
        if train.shape[0] < n_lists:
            train = self.db
        self.cent = kmeans(train, n_lists, iters, seed, balance=balance, init=init)
        self.labels = assign(self.db, self.cent)
        order = torch.argsort(self.labels, stable=True)
        self.order = order
        counts = torch.bincount(self.labels, minlength=n_lists)
        self.offsets = torch.zeros(n_lists + 1, dtype=torch.int64)
        self.offsets[1:] = torch.cumsum(counts, 0)

    def probes(self, q: torch.Tensor, n_probes: int) -> torch.Tensor:
        q = q.to(torch.float32)
        if self.metric in ("sqeuclidean", "l2", "L2"):
            score = (self.cent * self.cent).sum(1)[None, :] - 2.0 * (q @ self.cent.T)
        else:
            score = -(q @ self.cent.T)
        return torch.topk(score, min(n_probes, self.cent.shape[0]), dim=1, largest=False).indices

    def search(self, q: torch.Tensor, k: int, n_probes: int = 20) -> Tuple[torch.Tensor, torch.Tensor]:
        q = q.to(torch.float32)
        pr = self.probes(q, n_probes)
        out_d = torch.full((q.shape[0], k), float("inf"))
        out_i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
        l2 = self.metric in ("sqeuclidean", "l2", "L2")
        for qi in range(q.shape[0]):
            rows = torch.cat([self.order[self.offsets[l]:self.offsets[l + 1]] for l in pr[qi].tolist()])
            if rows.numel() == 0:
                continue
            xs = self.db[rows]
            if l2:
                sc = ((xs - q[qi][None, :]) ** 2).sum(1)
            else:
                sc = -(xs @ q[qi])
            kk = min(k, rows.numel())
            d, pos = torch.topk(sc, kk, largest=False, sorted=True)
            out_d[qi, :kk] = d if l2 else -d
            out_i[qi, :kk] = rows[pos]
        if not l2:
            out_d[out_i < 0] = float("-inf")
        return out_d, out_i


    def search_many(self, q: torch.Tensor, k: int, n_probes_list, block: int = 256):
        """Batched form of ``search`` for recall sweeps: {n_probes: ids [Q, k]} for several probe
        counts at once.  Same semantics (exact distances to every row of the probed lists), computed
        as one dense distance block per query chunk with rows outside the probed lists masked out."""
        q = q.to(torch.float32)
        l2 = self.metric in ("sqeuclidean", "l2", "L2")
        n_lists = self.cent.shape[0]
        pmax = min(max(n_probes_list), n_lists)
        pr_all = self.probes(q, pmax)
        out = {p: torch.full((q.shape[0], k), -1, dtype=torch.int64) for p in n_probes_list}
        xn = (self.db * self.db).sum(1) if l2 else None
        for s in range(0, q.shape[0], block):
            qb = q[s:s + block]
            sc = -(qb @ self.db.T)
            if l2:
                sc = 2.0 * sc + xn[None, :]          # ||x||^2 - 2 q.x (the per-query constant is dropped)
            for p in n_probes_list:
                member = torch.zeros((qb.shape[0], n_lists), dtype=torch.bool)
                member.scatter_(1, pr_all[s:s + block, :min(p, n_lists)], True)
                masked = sc.masked_fill(~member[:, self.labels], float("inf"))
                d, idx = torch.topk(masked, k, dim=1, largest=False, sorted=True)
                idx = idx.masked_fill(torch.isinf(d), -1)
                out[p][s:s + block] = idx
        return out


class IvfPqOracle(IvfFlatOracle):
    def __init__(self, db: torch.Tensor, n_lists: int, pq_dim: int, metric: str = "sqeuclidean",
                 iters: int = 20, train_fraction: float = 0.5, seed: int = 0, pq_iters: int = 10,
                 n_codes: int = 256, balance: bool = True, init=None, pq_init_rows=None):
        super().__init__(db, n_lists, metric, iters, train_fraction, seed, balance=balance, init=init)
        d = self.db.shape[1]
        assert d % pq_dim == 0
        self.pq_dim, self.dsub = pq_dim, d // pq_dim
        res = self.db - self.cent[self.labels]
        n = res.shape[0]
        stride = max(1, n // 131072)
        tr = res[::stride]
        self.codebooks = torch.stack([
            kmeans(tr[:, m * self.dsub:(m + 1) * self.dsub], n_codes, pq_iters, seed + 31 * (m + 1),
                   balance=balance,
                   init=None if pq_init_rows is None else tr[pq_init_rows, m * self.dsub:(m + 1) * self.dsub])
            for m in range(pq_dim)])  # [M, n_codes (256), dsub]
        codes = torch.empty((n, pq_dim), dtype=torch.int64)
        for m in range(pq_dim):
            codes[:, m] = assign(res[:, m * self.dsub:(m + 1) * self.dsub], self.codebooks[m])
        self.codes = codes

    def search(self, q: torch.Tensor, k: int, n_probes: int = 20, refine_ratio: int = 1):
        """ADC search; ``refine_ratio > 1`` re-ranks the best refine_ratio*k ADC candidates with
        exact distances to the original rows (cuVS ``refine`` semantics)."""
        if refine_ratio > 1:
            _, cand = self.search(q, k * refine_ratio, n_probes, 1)
            return self._refine(q.to(torch.float32), cand, k)
        q = q.to(torch.float32)
        pr = self.probes(q, n_probes)
        out_d = torch.full((q.shape[0], k), float("inf"))
        out_i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
        l2 = self.metric in ("sqeuclidean", "l2", "L2")
        M, ds = self.pq_dim, self.dsub
        for qi in range(q.shape[0]):
            cand_s, cand_r = [], []
            for l in pr[qi].tolist():
                rows = self.order[self.offsets[l]:self.offsets[l + 1]]
                if rows.numel() == 0:
                    continue
                if l2:
                    rq = (q[qi] - self.cent[l]).view(M, 1, ds)
                    lut = ((rq - self.codebooks) ** 2).sum(2)          # [M, 256]
                    bias = 0.0
                else:
                    lut = -(self.codebooks * q[qi].view(M, 1, ds)).sum(2)
                    bias = -float(q[qi] @ self.cent[l])
                sc = bias + lut[torch.arange(M)[None, :], self.codes[rows]].sum(1)
                cand_s.append(sc)
                cand_r.append(rows)
            if not cand_s:
                continue
            sc = torch.cat(cand_s)
            rows = torch.cat(cand_r)
            kk = min(k, rows.numel())
            d, pos = torch.topk(sc, kk, largest=False, sorted=True)
            out_d[qi, :kk] = d if l2 else -d
            out_i[qi, :kk] = rows[pos]
        return out_d, out_i

    def _refine(self, q, cand, k):
        l2 = self.metric in ("sqeuclidean", "l2", "L2")
        out_d = torch.full((q.shape[0], k), float("inf") if l2 else float("-inf"))
        out_i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
        for qi in range(q.shape[0]):
            rows = cand[qi][cand[qi] >= 0]
            if rows.numel() == 0:
                continue
            xs = self.db[rows]
            sc = ((xs - q[qi][None, :]) ** 2).sum(1) if l2 else -(xs @ q[qi])
            kk = min(k, rows.numel())
            d, pos = torch.topk(sc, kk, largest=False, sorted=True)
            out_d[qi, :kk] = d if l2 else -d
            out_i[qi, :kk] = rows[pos]
        return out_d, out_i


def recall(ids: torch.Tensor, truth: torch.Tensor) -> float:
    hits = 0
    for a, b in zip(ids.tolist(), truth.tolist()):
        hits += len(set(a) & set(x for x in b if x >= 0))
    denom = int((truth >= 0).sum())
    return hits / float(max(denom, 1))
