"""IVF recall parity on HARD data (VERDICT r1 "harden IVF parity"): iid Gaussian and unit-norm
corpora have no cluster structure, so recall depends on the index actually scanning the same
share of each query's true neighbourhood as the restated reference-semantics index - an index that
merely finds "the right cluster" scores nothing here.  Shapes follow BASELINE configs C3 (768-d
fp16 IVF-Flat) and C4 (128-d fp16 IVF-PQ, M=64) at the rows-per-list ratio of the full-size runs;
queries are drawn INDEPENDENTLY from the data distribution (never perturbed database rows).

Criterion (north star): recall@10 against exact ground truth equals the oracle index's recall at
identical n_lists / n_probes, |diff| <= 0.02, for n_probes in {1, 8, 32, 128}.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 0.02   # recall@10 difference GPU index vs oracle index at the same n_lists / n_probes


def corpus(kind, n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    if kind == "unit":
        x = torch.nn.functional.normalize(x, dim=1)
    return x


SEEDS = (0, 1, 2)


@pytest.mark.parametrize("kind,metric", [("gauss", "sqeuclidean"), ("unit", "sqeuclidean"),
                                         ("unit", "inner_product")])
def test_ivf_flat_recall_parity_on_structureless_data_c3_shape(b2, kind, metric):
    """k-means on structureless rows is chaotic: the cluster-size spread of two runs of the SAME
    implementation differs by ~10 % between seeds (GPU: 60.5 / 69.2 / 77.9, oracle: 65.4 / 65.9 /
    68.0 on this corpus, profiles/r2_kmeans_seed_spread.txt), which moves recall@10 at 128 probes by
    +-0.02.  The parity statement is therefore about the MEAN over three independently seeded
    indexes on each side, same n_lists / n_probes."""
    from oracle.exact import exact_knn
    from oracle.ivf import IvfFlatOracle, recall
    n, d, nlist, k, nq = 196_608, 768, 1024, 10, 1000      # 192 rows per list; C3 has 2441
    x = corpus(kind, n, d, 31).to(torch.float16)
    q = corpus(kind, nq, d, 32).to(torch.float16)           # independent draws, not database rows
    _, truth = exact_knn(x.float(), q.float(), k, metric)
    probes = [1, 8, 32, 128]
    gpu = {p: [] for p in probes}
    ref = {p: [] for p in probes}
    xg, qg = x.cuda(), q.cuda()
    for seed in SEEDS:
        ix = b2.NativeIndex.ivf_flat(xg, nlist, metric=metric, kmeans_iters=8, seed=seed)
        for p in probes:
            _, gi = ix.search(qg, k, n_probes=p)
            gpu[p].append(recall(gi.cpu(), truth))
        ix.destroy()
        o_ids = IvfFlatOracle(x.float(), nlist, metric, iters=8, seed=seed).search_many(q.float(), k, probes)
        for p in probes:
            ref[p].append(recall(o_ids[p], truth))
    mean = lambda v: sum(v) / len(v)
    report = {p: (round(mean(gpu[p]), 4), round(mean(ref[p]), 4)) for p in probes}
    for p, (r_gpu, r_ref) in report.items():
        assert abs(r_gpu - r_ref) <= TOL, (report, gpu, ref)
    # structureless data: recall must actually move with the probe count (no trivially-1.0 corpus)
    assert report[1][0] < report[8][0] < report[32][0] < report[128][0] < 0.9, report
    assert report[128][0] > 3 * report[8][0], report


def test_ivf_pq_recall_parity_on_structureless_data_c4_shape(b2):
    from oracle.exact import exact_knn
    from oracle.ivf import IvfPqOracle, recall
    n, d, nlist, m, k, nq = 131_072, 128, 512, 64, 10, 600   # 256 rows per list, M = 64 (dsub 2) as C4
    x = corpus("gauss", n, d, 41).to(torch.float16)
    q = corpus("gauss", nq, d, 42).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, m, kmeans_iters=8)
    _, truth = exact_knn(x.float(), q.float(), k)
    oracle = IvfPqOracle(x.float(), nlist, m, iters=8, pq_iters=8)
    report = {}
    for p, rr in ((8, 1), (64, 1), (64, 4), (64, 20)):   # 20 x 10 = 200 candidates: beyond the old 128 clamp
        _, gi = ix.search(q.cuda(), k, n_probes=p, refine_ratio=rr)
        _, oi = oracle.search(q.float(), k, n_probes=p, refine_ratio=rr)
        report[(p, rr)] = (round(recall(gi.cpu(), truth), 4), round(recall(oi, truth), 4))
    for key, (r_gpu, r_ref) in report.items():
        assert abs(r_gpu - r_ref) <= TOL + 0.01, report     # + PQ codebook training noise
    # M = 64 on 128-d (8 bits per 2 dims) is a fine quantizer: refine only adds a little here
    assert report[(64, 4)][0] >= report[(64, 1)][0] - 0.002, report
    assert report[(64, 20)][0] >= report[(64, 4)][0] - 0.002, report
    assert report[(8, 1)][0] < 0.5 * report[(64, 1)][0], report


def test_ivf_pq_default_params_build_at_dim_768(b2, ):
    """IndexBuildConfig('ivf_pq', {}) must build at the headline dims: the default pq_dim follows the
    reference's min(64, dim // 4) snapped to a supported divisor."""
    grm = b2.GPUResourceManager(devices=[0])
    ibc = b2.IndexBuildingCoordinator(grm)
    x = corpus("gauss", 20_000, 768, 5).to(torch.float16).cuda()
    part = b2.EmbeddingPart(0, x, 0, x.shape[0])
    res = ibc._build_single_index(part, b2.IndexBuildConfig("ivf_pq", {"n_lists": 32}, parallel_build=False,
                                                            max_retries=0))
    assert res.success, res.error_message
    inf = res.index.info()
    assert inf.pq_dim == 96 and inf.dim == 768
    d, i = res.index.search(x[:16], 5, n_probes=8, refine_ratio=4)
    assert (i[:, 0].cpu() == torch.arange(16)).all()


def test_cosine_metric_equals_sklearn_cosine_golden(b2, golden_dir):
    """B2VS_METRIC_COSINE = scikit-learn NearestNeighbors(metric='cosine', algorithm='brute'), the
    reference's CPU baseline (VectorSearch_QuestionRetrieval.ipynb:L878): golden from sklearn."""
    import os
    import numpy as np
    g = np.load(os.path.join(golden_dir, "ivf.npz"))
    x, q, k = torch.from_numpy(g["x"]), torch.from_numpy(g["q"]), int(g["k"])
    ix = b2.NativeIndex.flat(x.cuda(), metric="cosine", id_offset=0)
    assert not ix.descending and ix.info().metric == 2
    d, i = ix.search(q.cuda(), k)
    assert (i.cpu().numpy() == g["cos_i"]).mean() > 0.99
    np.testing.assert_allclose(d.cpu().numpy(), g["cos_d"], atol=3e-5)
    assert (d[:, 1:] >= d[:, :-1]).all()
    # IVF-Flat with every list probed: the same answer up to its 16-bit list storage (fp32 rows are
    # kept as bf16 in the lists: ~4e-3 relative on similarities close to 1)
    iv = b2.NativeIndex.ivf_flat(x.cuda(), 16, metric="cosine", kmeans_iters=4)
    assert iv.info().metric == 2
    d2, i2 = iv.search(q.cuda(), k, n_probes=16)
    inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i2.cpu(), torch.from_numpy(g["cos_i"])))
    assert inter >= 0.8 * g["cos_i"].size, inter
    np.testing.assert_allclose(d2.cpu().numpy(), g["cos_d"], atol=8e-3)
    assert (d2[:, 1:] >= d2[:, :-1]).all()


def test_query_dim_mismatch_and_bad_out_buffers_are_rejected(b2):
    """a narrower query matrix would be read out of bounds, a wider one silently gives
    wrong neighbours; caller-supplied out tensors are validated before any device write."""
    x = corpus("gauss", 4096, 64, 1).to(torch.float16).cuda()
    ix = b2.NativeIndex.flat(x)
    for bad in (32, 128):
        with pytest.raises(ValueError, match="dim"):
            ix.search(torch.zeros(4, bad, dtype=torch.float16, device="cuda"), 3)
    import ctypes
    n = b2._native
    qz = torch.zeros(4, 32, dtype=torch.float16, device="cuda")
    d = torch.empty((4, 3), dtype=torch.float32, device="cuda")
    i = torch.empty((4, 3), dtype=torch.int64, device="cuda")
    rc = n.lib().b2vs_search(ix._h, qz.data_ptr(), n.F16, 4, 32, 3, None, d.data_ptr(), i.data_ptr(), None)
    assert rc == -1 and b"dim" in n.lib().b2vs_last_error()
    q = x[:4].contiguous()
    with pytest.raises(ValueError, match="out ids"):
        ix.search(q, 3, out=(d, torch.empty((4, 3), dtype=torch.int32, device="cuda")))
    with pytest.raises(ValueError, match="out distances"):
        ix.search(q, 3, out=(torch.empty((4, 2), dtype=torch.float32, device="cuda"), i))
    with pytest.raises(ValueError, match="out distances"):
        ix.search(q, 3, out=(torch.empty((4, 3), dtype=torch.float32), i))


def test_refine_without_source_rows_is_an_error(b2, tmp_path):
    x = corpus("gauss", 8192, 64, 2).to(torch.float16).cuda()
    ix = b2.NativeIndex.ivf_pq(x, 16, 32, kmeans_iters=4)
    path = str(tmp_path / "pq.b2vs")
    ix.save(path)
    bare = b2.NativeIndex.load(path, "cuda:0")
    with pytest.raises(ValueError, match="source rows"):
        bare.search(x[:8].contiguous(), 5, n_probes=4, refine_ratio=4)
    import ctypes
    n = b2._native
    d = torch.empty((8, 5), dtype=torch.float32, device="cuda")
    i = torch.empty((8, 5), dtype=torch.int64, device="cuda")
    sp = n.SearchParams(4, 4, 0, 0)
    rc = n.lib().b2vs_search(bare._h, x.data_ptr(), n.F16, 8, 64, 5, ctypes.byref(sp), d.data_ptr(),
                             i.data_ptr(), None)
    assert rc == -1 and b"rows_for_refine" in n.lib().b2vs_last_error()
    full = b2.NativeIndex.load(path, "cuda:0", rows=x)
    _, ids = full.search(x[:8].contiguous(), 5, n_probes=4, refine_ratio=4)
    assert (ids[:, 0].cpu() == torch.arange(8)).all()


def test_corrupt_index_files_are_rejected(b2, tmp_path):
    """b2vs_index_load re-derives every section size from the header scalars."""
    x = corpus("gauss", 8192, 64, 3).to(torch.float16).cuda()
    ix = b2.NativeIndex.ivf_flat(x, 16, kmeans_iters=4)
    path = str(tmp_path / "flat.b2vs")
    ix.save(path)
    raw = bytearray(open(path, "rb").read())
    import struct
    cases = {}
    t = bytearray(raw); struct.pack_into("<i", t, 8 + 4 * 4, -5); cases["negative n_lists"] = t
    t = bytearray(raw); struct.pack_into("<i", t, 8 + 4 * 4, 1 << 30); cases["huge n_lists"] = t
    t = bytearray(raw); struct.pack_into("<i", t, 8 + 3 * 4, 4096); cases["dim"] = t
    t = bytearray(raw); struct.pack_into("<i", t, 8, 7); cases["kind"] = t
    cases["truncated"] = raw[: len(raw) // 2]
    cases["trailing bytes"] = raw + b"xx"
    for name, blob in cases.items():
        p = str(tmp_path / "bad.b2vs")
        open(p, "wb").write(bytes(blob))
        with pytest.raises(RuntimeError, match="b2vs_index_load failed"):
            b2.NativeIndex.load(p, "cuda:0")
    again = b2.NativeIndex.load(path, "cuda:0")     # the untouched file still loads
    _, ids = again.search(x[:8].contiguous(), 3, n_probes=16)
    assert (ids[:, 0].cpu() == torch.arange(8)).all()
