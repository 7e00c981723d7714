"""Pins the CPU oracle against the reference's own known answers and golden fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import exact, merge
from oracle.ivf import IvfFlatOracle, IvfPqOracle, recall


# ---- merge: known answers of Attempt_1/test_search_result_aggregator.py:308-358
def test_merge_single_gpu_known_answer():
    d = [np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float32)]
    i = [np.array([[10, 20, 30], [40, 50, 60]], np.int64)]
    fd, fi = merge.merge_topk(d, i, 2)
    np.testing.assert_array_equal(fd, [[1.0, 2.0], [4.0, 5.0]])
    np.testing.assert_array_equal(fi, [[10, 20], [40, 50]])


def test_merge_two_gpu_known_answer():
    d = [np.array([[2.0, 4.0], [6.0, 8.0]], np.float32), np.array([[1.0, 3.0], [5.0, 7.0]], np.float32)]
    i = [np.array([[20, 40], [60, 80]], np.int64), np.array([[10, 30], [50, 70]], np.int64)]
    fd, fi = merge.merge_topk(d, i, 3)
    np.testing.assert_array_equal(fd, [[1.0, 2.0, 3.0], [5.0, 6.0, 7.0]])
    np.testing.assert_array_equal(fi, [[10, 20, 30], [50, 60, 70]])


# ---- partitioning: reference known answers + fixtures generated from the reference itself
def test_partition_known_answers():
    assert merge.partition_even(300, 3) == [(0, 0, 100), (1, 100, 200), (2, 200, 300)]
    assert merge.partition_even(301, 3) == [(0, 0, 101), (1, 101, 201), (2, 201, 301)]


def test_partition_matches_reference_outputs(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "partition.json")))
    n_checked = 0
    for c in cases:
        if c["strategy"] != "even" or c["n"] < c["g"]:
            continue  # the reference emits empty ranges when n < g; both are degenerate
        assert [list(t) for t in merge.partition_even(c["n"], c["g"])] == c["out"], c
        n_checked += 1
    assert n_checked > 40


def test_uneven_shard_id_offsets():
    # 750k rows over 2 shards load as 375000 + 375000; 750001 -> 375001 + 375000 (bug 1)
    parts = merge.partition_even(750001, 2)
    assert parts[1][1] == 375001
    ids = merge.global_ids(np.array([[0, 5, -1]]), parts[1][1])
    np.testing.assert_array_equal(ids, [[375001, 375006, -1]])


# ---- exact search: sklearn brute (the reference's CPU baseline) and golden vectors
def test_exact_knn_matches_golden_l2(golden_dir):
    g = np.load(os.path.join(golden_dir, "knn_l2.npz"))
    d, i = exact.exact_knn(torch.from_numpy(g["db"]), torch.from_numpy(g["q"]), 10, "sqeuclidean")
    assert (i.numpy() == g["i"]).mean() > 0.999
    np.testing.assert_allclose(d.numpy(), g["d"], rtol=2e-4, atol=2e-4)


def test_exact_knn_matches_golden_ip(golden_dir):
    g = np.load(os.path.join(golden_dir, "knn_ip.npz"))
    d, i = exact.exact_knn(torch.from_numpy(g["db"]), torch.from_numpy(g["q"]), 10, "inner_product")
    assert (i.numpy() == g["i"]).mean() > 0.999
    np.testing.assert_allclose(d.numpy(), g["d"], rtol=1e-4, atol=1e-5)


def test_exact_knn_on_reference_sample_embeddings(golden_dir):
    g = np.load(os.path.join(golden_dir, "sample_emb.npz"))
    emb = torch.from_numpy(g["emb"])
    assert emb.shape == (10, 384)
    np.testing.assert_allclose(np.linalg.norm(g["emb"], axis=1), 1.0, atol=1e-4)
    d, i = exact.exact_knn(emb, emb, 5, "inner_product")
    np.testing.assert_array_equal(i.numpy(), g["i"])
    assert (i[:, 0] == torch.arange(10)).all()  # every unit-norm row is its own best match


def test_exact_knn_agrees_with_sklearn_brute_live():
    g = torch.Generator().manual_seed(3)
    db, q = torch.randn(3000, 48, generator=g), torch.randn(25, 48, generator=g)
    d, i = exact.exact_knn(db, q, 7)
    d2, i2 = exact.sklearn_brute_knn(db.numpy(), q.numpy(), 7)
    assert (i.numpy() == i2).mean() > 0.999
    np.testing.assert_allclose(d.numpy(), d2, rtol=1e-3, atol=1e-3)


def test_exact_knn_k_larger_than_n_and_blocks():
    g = torch.Generator().manual_seed(4)
    db, q = torch.randn(30, 16, generator=g), torch.randn(4, 16, generator=g)
    d, i = exact.exact_knn(db, q, 40, block=7)
    assert (i[:, 30:] == -1).all() and torch.isinf(d[:, 30:]).all()
    assert sorted(i[0, :30].tolist()) == list(range(30))
    d1, i1 = exact.exact_knn(db, q, 5, block=7)
    d2, i2 = exact.exact_knn(db, q, 5, block=1000)
    assert (i1 == i2).all()


def test_parity_report_flags_wrong_results():
    g = torch.Generator().manual_seed(5)
    db, q = torch.randn(500, 16, generator=g), torch.randn(6, 16, generator=g)
    d, i = exact.exact_knn(db, q, 5)
    assert exact.topk_parity_report(d, i, db, q, 5)["ok"]
    bad = i.clone()
    bad[0, 0] = int(exact.pairwise_f64(db, q)[0].argmax())
    assert not exact.topk_parity_report(d, bad, db, q, 5)["ok"]


# ---- IVF oracle sanity (semantics: full probe == exact; recall grows with n_probes)
def _clustered(n, d, c, seed):
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(c, d, generator=g)
    return cent[torch.randint(0, c, (n,), generator=g)] + 0.5 * torch.randn(n, d, generator=g)


def test_ivf_flat_oracle_full_probe_is_exact():
    x = _clustered(3000, 24, 20, 1)
    q = x[:40] + 0.05
    o = IvfFlatOracle(x, 16, iters=5)
    td, ti = exact.exact_knn(x, q, 5)
    d, i = o.search(q, 5, n_probes=16)
    assert recall(i, ti) == 1.0
    d1, i1 = o.search(q, 5, n_probes=1)
    d4, i4 = o.search(q, 5, n_probes=4)
    assert recall(i1, ti) <= recall(i4, ti) <= 1.0
    assert int((o.offsets[1:] - o.offsets[:-1]).sum()) == 3000


def test_ivf_pq_oracle_recall_reasonable():
    x = _clustered(4000, 32, 20, 2)
    q = x[:30] + 0.05
    o = IvfPqOracle(x, 8, pq_dim=16, iters=5, pq_iters=4)
    td, ti = exact.exact_knn(x, q, 10)
    d, i = o.search(q, 10, n_probes=8)
    assert recall(i, ti) > 0.7


# ---- encoder hand-off: the reference's own last_token_pool + F.normalize outputs
#      (tests/golden/encode.npz, generated by importing generate_embeddings.py:11-21)
ENCODE_CASES = ["left_padded", "right_padded", "zero_row", "all_ones", "single", "zero_vector"]


@pytest.mark.parametrize("case", ENCODE_CASES)
def test_encode_oracle_matches_reference_golden(case, golden_dir):
    from oracle import encode
    g = np.load(os.path.join(golden_dir, "encode.npz"))
    h, m = g[case + "_hidden"], g[case + "_mask"]
    np.testing.assert_array_equal(encode.last_token_pool(h, m), g[case + "_pooled"])   # a gather: bit-exact
    np.testing.assert_allclose(encode.pool_normalize(h, m), g[case + "_normalized"], rtol=2e-6, atol=1e-12)
    assert encode.pool_normalize(h, m, normalize=False).dtype == np.float32


def test_encode_oracle_mean_pool_definition():
    from oracle import encode
    h = np.arange(2 * 3 * 2, dtype=np.float32).reshape(2, 3, 2)
    m = np.array([[1, 1, 0], [0, 0, 0]], np.int64)
    out = encode.mean_pool(h, m)
    np.testing.assert_allclose(out[0], (h[0, 0] + h[0, 1]) / 2)
    np.testing.assert_array_equal(out[1], [0.0, 0.0])          # 0 / clamp(0, 1e-9)
    np.testing.assert_allclose(encode.mean_pool(h, None), h.mean(axis=1))


# ---- IVF restatement anchored on scikit-learn (tests/golden/ivf.npz, make_golden_ivf.py) ----------
@pytest.fixture(scope="module")
def ivf_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ivf.npz"))


def test_oracle_kmeans_equals_sklearn_lloyd_from_the_same_init(ivf_gold):
    from oracle.ivf import assign, kmeans
    g = ivf_gold
    x = torch.from_numpy(g["x"])
    cent = kmeans(x, len(g["init_rows"]), iters=int(g["iters"]), balance=False,
                  init=x[torch.from_numpy(g["init_rows"])])
    np.testing.assert_allclose(cent.numpy(), g["centroids"], rtol=2e-4, atol=2e-4)
    # list assignment == KMeans.predict
    lab = assign(x, torch.from_numpy(g["centroids"]))
    assert (lab.numpy() == g["labels"]).mean() > 0.999   # fp32 vs fp64 near-ties only


def test_oracle_probes_and_list_scan_equal_sklearn(ivf_gold):
    g = ivf_gold
    x, q = torch.from_numpy(g["x"]), torch.from_numpy(g["q"])
    o = IvfFlatOracle(x, len(g["init_rows"]), iters=int(g["iters"]), train_fraction=1.0, balance=False,
                      init=x[torch.from_numpy(g["init_rows"])])
    pr = o.probes(q, int(g["n_probes"]))
    assert np.array_equal(pr.numpy(), g["probes"])
    d, i = o.search(q, int(g["k"]), n_probes=int(g["n_probes"]))
    assert np.array_equal(i.numpy(), g["scan_i"])
    np.testing.assert_allclose(d.numpy(), g["scan_d"], rtol=1e-4, atol=1e-4)


def test_oracle_residual_pq_equals_sklearn_codebooks_and_adc(ivf_gold):
    g = ivf_gold
    x, q = torch.from_numpy(g["x"]), torch.from_numpy(g["q"])
    o = IvfPqOracle(x, len(g["init_rows"]), int(g["pq_M"]), iters=int(g["iters"]), train_fraction=1.0,
                    pq_iters=int(g["pq_iters"]), n_codes=int(g["pq_ncode"]), balance=False,
                    init=x[torch.from_numpy(g["init_rows"])], pq_init_rows=torch.from_numpy(g["pq_init_rows"]))
    np.testing.assert_allclose(o.codebooks.numpy(), g["pq_codebooks"], rtol=5e-4, atol=5e-4)
    assert (o.codes.numpy() == g["pq_codes"]).mean() > 0.995
    # ADC over the nearest list only == per-subspace squared distances to the chosen codebook rows
    # (16 codes x 4 subspaces: many rows share a code word, so ids tie; the sorted distances do not)
    d, i = o.search(q, int(g["k"]), n_probes=1)
    np.testing.assert_allclose(d.numpy(), g["adc_d"], rtol=2e-3, atol=2e-3)
    strict = np.ones_like(g["adc_d"], bool)                      # positions with a unique distance
    strict[:, 1:] &= np.abs(np.diff(g["adc_d"], axis=1)) > 1e-3
    strict[:, :-1] &= np.abs(np.diff(g["adc_d"], axis=1)) > 1e-3
    assert (i.numpy()[strict] == g["adc_rows"][strict]).mean() > 0.97


def test_oracle_refine_candidates_are_not_clamped():
    """k * refine_ratio > 128 candidates are re-ranked (the GPU path does not clamp either)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3000, 16, generator=g)
    q = torch.randn(8, 16, generator=g)
    o = IvfPqOracle(x, 8, 4, iters=4, pq_iters=3)
    seen = {}
    orig = o._refine
    o._refine = lambda qq, cand, k: seen.setdefault("n", cand.shape[1]) and orig(qq, cand, k)
    o.search(q, 40, n_probes=8, refine_ratio=5)
    assert seen["n"] == 200


def test_cosine_golden_is_one_minus_normalised_inner_product(ivf_gold):
    """B2VS_METRIC_COSINE semantics = scikit-learn metric='cosine' (the reference's CPU baseline,
    VectorSearch_QuestionRetrieval.ipynb:L878): checked against the exact oracle on unit rows."""
    g = ivf_gold
    x = torch.nn.functional.normalize(torch.from_numpy(g["x"]), dim=1)
    q = torch.nn.functional.normalize(torch.from_numpy(g["q"]), dim=1)
    d, i = exact.exact_knn(x, q, int(g["k"]), "inner_product")
    assert (i.numpy() == g["cos_i"]).mean() > 0.99
    np.testing.assert_allclose(1.0 - d.numpy(), g["cos_d"], atol=2e-5)


def test_oracle_search_many_equals_search():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4000, 24, generator=g)
    q = torch.randn(30, 24, generator=g)
    for metric in ("sqeuclidean", "inner_product"):
        o = IvfFlatOracle(x, 32, metric, iters=5)
        many = o.search_many(q, 7, [1, 4, 32])
        for p in (1, 4, 32):
            _, i = o.search(q, 7, n_probes=p)
            assert torch.equal(many[p], i), (metric, p)
