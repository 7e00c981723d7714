"""Generates tests/golden/ivf.npz: an EXTERNAL anchor for the IVF restatement in oracle/ivf.py.

The reference pins nothing at its cuVS / FAISS boundary (every test mocks the ANN calls, SURVEY
§8c), so the oracle's IVF semantics are anchored on scikit-learn - the library the reference's own
CPU baseline uses (Attempt_1/VectorSearch_QuestionRetrieval.ipynb:L878) - at each step:

  k-means     sklearn.cluster.KMeans(init=<fixed rows>, n_init=1, algorithm='lloyd', tol=0,
              max_iter=T): centroids after T Lloyd iterations from the same initial rows
  assignment  KMeans.predict (nearest centroid)
  probes      NearestNeighbors(brute) over the centroids: the n_probes nearest lists per query
  list scan   NearestNeighbors(brute) over the rows of the probed lists only: what an exact
              scan of those lists must return (ids + squared distances)
  PQ          per-subspace KMeans(256 -> 16 here) on residuals, codes = predict, ADC distance
              = sum of per-subspace squared distances to the chosen codebook rows

Also holds a cosine case: NearestNeighbors(metric='cosine', algorithm='brute') - the exact call of
the reference's CPU baseline - for the B2VS_METRIC_COSINE path.

Run in the build container (scikit-learn 1.9); the .npz travels, nothing here runs on the GPU box.
"""
import os

import numpy as np
from sklearn.cluster import KMeans
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))


def lloyd(x, init, iters):
    km = KMeans(n_clusters=init.shape[0], init=init, n_init=1, max_iter=iters, tol=0.0, algorithm="lloyd")
    km.fit(x)
    return km


def main():
    rng = np.random.default_rng(20261018)
    n, d, n_lists, iters, nq, n_probes, k = 3000, 16, 24, 6, 25, 4, 5
    # blobs, so no cluster empties and Lloyd converges to well-separated centres
    centres = rng.standard_normal((40, d)).astype(np.float32) * 3.0
    x = (centres[rng.integers(0, 40, n)] + 0.7 * rng.standard_normal((n, d))).astype(np.float32)
    q = (centres[rng.integers(0, 40, nq)] + 0.7 * rng.standard_normal((nq, d))).astype(np.float32)
    init_rows = rng.choice(n, n_lists, replace=False)
    km = lloyd(x.astype(np.float64), x[init_rows].astype(np.float64), iters)
    assert km.n_iter_ == iters, km.n_iter_
    cent = km.cluster_centers_
    labels = km.predict(x.astype(np.float64))
    assert len(np.unique(labels)) == n_lists
    # probes: nearest centroids (euclidean order == squared-euclidean order)
    pr = NearestNeighbors(n_neighbors=n_probes, algorithm="brute").fit(cent).kneighbors(
        q.astype(np.float64), return_distance=False)
    # exact scan of the probed lists
    scan_i = np.full((nq, k), -1, np.int64)
    scan_d = np.full((nq, k), np.inf, np.float64)
    for qi in range(nq):
        rows = np.nonzero(np.isin(labels, pr[qi]))[0]
        kk = min(k, len(rows))
        dd, ii = NearestNeighbors(n_neighbors=kk, algorithm="brute").fit(x[rows].astype(np.float64)).kneighbors(
            q[qi:qi + 1].astype(np.float64))
        scan_i[qi, :kk] = rows[ii[0]]
        scan_d[qi, :kk] = dd[0] ** 2
    # residual PQ: M sub-quantizers x 16 codes (small so the fixture stays tiny)
    M, ncode, pq_iters = 4, 16, 5
    ds = d // M
    res = x.astype(np.float64) - cent[labels]
    cb = np.zeros((M, ncode, ds))
    codes = np.zeros((n, M), np.int64)
    pq_init_rows = rng.choice(n, ncode, replace=False)
    for m in range(M):
        sub = res[:, m * ds:(m + 1) * ds]
        kmm = lloyd(sub, sub[pq_init_rows], pq_iters)
        cb[m] = kmm.cluster_centers_
        codes[:, m] = kmm.predict(sub)
    # ADC distances of query 0..nq-1 to every row of their FIRST probed list
    adc_rows, adc_d = [], []
    for qi in range(nq):
        l = pr[qi, 0]
        rows = np.nonzero(labels == l)[0]
        rq = q[qi].astype(np.float64) - cent[l]
        dist = np.zeros(len(rows))
        for m in range(M):
            dist += ((rq[m * ds:(m + 1) * ds][None, :] - cb[m][codes[rows, m]]) ** 2).sum(1)
        order = np.argsort(dist, kind="stable")[:k]
        adc_rows.append(np.pad(rows[order], (0, k - len(order)), constant_values=-1))
        adc_d.append(np.pad(dist[order], (0, k - len(order)), constant_values=np.inf))
    # cosine: the reference's CPU baseline call
    cd, ci = NearestNeighbors(n_neighbors=k, metric="cosine", algorithm="brute").fit(x).kneighbors(q)
    np.savez_compressed(
        os.path.join(HERE, "ivf.npz"), x=x, q=q, init_rows=init_rows.astype(np.int64), iters=iters,
        centroids=cent.astype(np.float32), labels=labels.astype(np.int64), probes=pr.astype(np.int64),
        scan_i=scan_i, scan_d=scan_d.astype(np.float32), n_probes=n_probes, k=k,
        pq_M=M, pq_ncode=ncode, pq_iters=pq_iters, pq_init_rows=pq_init_rows.astype(np.int64),
        pq_codebooks=cb.astype(np.float32), pq_codes=codes, adc_rows=np.stack(adc_rows).astype(np.int64),
        adc_d=np.stack(adc_d).astype(np.float32), cos_d=cd.astype(np.float32), cos_i=ci.astype(np.int64))
    print("wrote ivf.npz")


if __name__ == "__main__":
    main()
