"""The selection rule behind FlatEngine::search_two_pass (csrc/flat.cu, csrc/merge.cu), restated in
numpy and checked on adversarial inputs: with T = the k-th smallest (chunk minimum, chunk index) pair
of a score row cut into 32-element chunks, the elements with (score, chunk) <= T
  * contain the k best elements in (score, index) order, and
  * number at most 32 * k - ties included.
The CUDA path relies on the second point for a candidate buffer that cannot overflow."""
import numpy as np
import pytest


def survivors(scores: np.ndarray, k: int):
    n = scores.shape[0]
    pad = (-n) % 32
    s = np.concatenate([scores, np.full(pad, np.inf, scores.dtype)]).reshape(-1, 32)
    mins = s.min(axis=1)
    order = np.lexsort((np.arange(mins.shape[0]), mins))        # by (minimum, chunk index)
    finite = np.isfinite(mins[order])
    if finite.sum() < k:                                         # fewer than k chunks: everything qualifies
        keep = np.isfinite(scores)
        return np.nonzero(keep)[0]
    t = order[k - 1]
    tau, tau_chunk = mins[t], t
    chunk = np.arange(n) // 32
    keep = (scores < tau) | ((scores == tau) & (chunk <= tau_chunk))
    return np.nonzero(keep)[0]


CASES = [
    ("gaussian", lambda g, n: g.standard_normal(n).astype(np.float32)),
    ("few distinct values", lambda g, n: g.integers(0, 4, n).astype(np.float32)),
    ("all equal", lambda g, n: np.full(n, 0.25, np.float32)),
    ("sorted ascending", lambda g, n: np.sort(g.standard_normal(n).astype(np.float32))),
    ("sorted descending", lambda g, n: -np.sort(-g.standard_normal(n).astype(np.float32))),
    ("one chunk holds every small value", lambda g, n: np.where(np.arange(n) // 32 == 3, -1.0, 1.0).astype(np.float32)
     + g.integers(0, 2, n).astype(np.float32) * 0.0),
]


@pytest.mark.parametrize("name,make", CASES)
@pytest.mark.parametrize("n,k", [(16384, 64), (4096, 32), (5003, 50), (1000, 128), (300, 2), (64, 64)])
def test_chunk_threshold_keeps_the_top_k_and_at_most_32k_elements(name, make, n, k):
    g = np.random.default_rng(hash((name, n, k)) % (2 ** 32))
    scores = make(g, n)
    keep = survivors(scores, k)
    assert keep.shape[0] <= 32 * k
    want = np.lexsort((np.arange(n), scores))[:min(k, n)]       # top-k in (score, index) order
    assert np.isin(want, keep).all()
