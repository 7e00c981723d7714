"""GPU parity of the exact-search path (K0 fused tcgen05 kernel + K8 merges) through the C ABI,
against the CPU oracle on identical seeded inputs.  Tolerance (north star): ids equal to an exact
fp32 search except for distance ties within 1e-3 relative; distances within 1e-3 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-3
DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def make(n, d, q, dtype, metric, seed=1234):
    g = torch.Generator().manual_seed(seed)
    db, qs = torch.randn(n, d, generator=g), torch.randn(q, d, generator=g)
    if metric == "inner_product":
        db = torch.nn.functional.normalize(db, dim=1)
        qs = torch.nn.functional.normalize(qs, dim=1)
    return db.to(DT[dtype]), qs.to(DT[dtype])


def check(b2, db, qs, k, metric, id_offset=0, **kw):
    from oracle.exact import topk_parity_report
    ix = b2.NativeIndex.flat(db.cuda(), metric=metric, id_offset=id_offset)
    d, i = ix.search(qs.cuda(), k, **kw)
    torch.cuda.synchronize()
    i = i.cpu()
    local = torch.where(i >= 0, i - id_offset, i)
    rep = topk_parity_report(d.cpu(), local, db.float(), qs.float(), k, metric, rtol=RTOL)
    assert rep["ok"], rep
    return d.cpu(), i, ix


CASES = [
    # n, d, q, k, dtype, metric, n_splits
    (1000, 64, 10, 5, "bf16", "sqeuclidean", 0),
    (256, 64, 128, 1, "bf16", "sqeuclidean", 0),
    (20000, 768, 300, 100, "bf16", "sqeuclidean", 0),
    (30000, 384, 257, 10, "fp16", "inner_product", 0),
    (5000, 96, 64, 10, "fp32", "sqeuclidean", 0),
    (3000, 100, 33, 7, "bf16", "sqeuclidean", 0),       # dim % 8 != 0: padded operand copy
    (50000, 128, 512, 32, "bf16", "sqeuclidean", 7),    # forced db splits
    (9000, 256, 130, 128, "fp16", "sqeuclidean", 0),    # k = fused maximum
    (256, 128, 40000, 1, "fp16", "sqeuclidean", 0),     # many items per CTA (k-means assign shape)
    (5000, 128, 20000, 10, "bf16", "inner_product", 0),
]


@pytest.mark.parametrize("n,d,q,k,dtype,metric,splits", CASES)
def test_exact_search_parity(b2, n, d, q, k, dtype, metric, splits):
    db, qs = make(n, d, q, dtype, metric)
    check(b2, db, qs, k, metric, id_offset=1000, n_splits=splits)


@pytest.mark.parametrize("group_flag", ["single", "pair"])
def test_both_kernel_variants_agree(b2, group_flag):
    """cta_group::1 and cta_group::2 kernels must give the same neighbours."""
    n = b2._native
    db, qs = make(40000, 256, 700, "bf16", "sqeuclidean")
    ix = n.NativeIndex.flat(db.cuda())
    import ctypes
    flag = 2 if group_flag == "single" else 4
    d = torch.empty((700, 50), dtype=torch.float32, device="cuda")
    i = torch.empty((700, 50), dtype=torch.int64, device="cuda")
    sp = n.SearchParams(0, 0, 0, flag)
    rc = n.lib().b2vs_search(ix._h, qs.cuda().data_ptr(), n.BF16, 700, 256, 50, ctypes.byref(sp),
                             d.data_ptr(), i.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, n.lib().b2vs_last_error()
    torch.cuda.synchronize()
    from oracle.exact import topk_parity_report
    assert topk_parity_report(d.cpu(), i.cpu(), db.float(), qs.float(), 50, rtol=RTOL)["ok"]


def test_config_c1_full_size(b2):
    """BASELINE configs[0]: exact inner product k=10, 100K x 384 fp32 unit-norm rows, 1K queries."""
    g = torch.Generator().manual_seed(4321)
    db = torch.nn.functional.normalize(torch.randn(100_000, 384, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(1000, 384, generator=g), dim=1)
    from oracle.exact import exact_knn
    ix = b2.NativeIndex.flat(db.cuda(), metric="inner_product")
    d, i = ix.search(qs.cuda(), 10)
    torch.cuda.synchronize()
    rd, ri = exact_knn(db, qs, 10, "inner_product")
    match = (i.cpu() == ri).float().mean().item()
    assert match > 0.999, match                      # ties within 1e-3 are the only allowed diffs
    np.testing.assert_allclose(d.cpu().numpy(), rd.numpy(), rtol=RTOL, atol=1e-4)


def test_k_larger_than_n_and_empty_tail(b2):
    db, qs = make(50, 64, 5, "bf16", "sqeuclidean")
    d, i, _ = check(b2, db, qs, 100, "sqeuclidean")
    assert (i[:, 50:] == -1).all() and torch.isinf(d[:, 50:]).all()


def test_golden_fixture_and_reference_sample_embeddings(b2, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "knn_l2.npz"))
    ix = b2.NativeIndex.flat(torch.from_numpy(g["db"]).cuda())
    d, i = ix.search(torch.from_numpy(g["q"]).cuda(), 10)
    assert (i.cpu().numpy() == g["i"]).mean() > 0.999
    np.testing.assert_allclose(d.cpu().numpy(), g["d"], rtol=RTOL, atol=1e-3)
    s = np.load(os.path.join(golden_dir, "sample_emb.npz"))
    emb = torch.from_numpy(s["emb"]).cuda()
    ix = b2.NativeIndex.flat(emb, metric="inner_product")
    d, i = ix.search(emb, 5)
    np.testing.assert_array_equal(i.cpu().numpy(), s["i"])


def test_search_host_matches_device_search(b2):
    db, qs = make(8000, 128, 300, "bf16", "sqeuclidean")
    ix = b2.NativeIndex.flat(db.cuda(), id_offset=5)
    d1, i1 = ix.search(qs.cuda(), 20)
    d2, i2 = ix.search_host(qs, 20)
    assert torch.equal(i1.cpu(), i2) and torch.allclose(d1.cpu(), d2)


def test_error_paths_return_codes_not_crashes(b2):
    db, qs = make(1000, 64, 4, "bf16", "sqeuclidean")
    ix = b2.NativeIndex.flat(db.cuda())
    with pytest.raises(RuntimeError, match="outside the supported range"):
        ix.search(qs.cuda(), 2049)
    with pytest.raises(ValueError, match="must live on a CUDA device"):
        ix.search(qs, 5)
    inf = ix.info()
    assert (inf.n_rows, inf.dim, inf.kind) == (1000, 64, 0)
    ix.destroy()
    with pytest.raises(RuntimeError, match="destroyed"):
        ix.search(qs.cuda(), 5)


# ---- K8 cross-shard merge on the GPU
def test_merge_topk_known_answers_on_gpu(b2):
    d = torch.tensor([[[2.0, 4.0], [6.0, 8.0]], [[1.0, 3.0], [5.0, 7.0]]]).cuda()
    i = torch.tensor([[[20, 40], [60, 80]], [[10, 30], [50, 70]]]).cuda()
    md, mi = b2.merge_topk(d, i, 3)
    assert mi.cpu().tolist() == [[10, 20, 30], [50, 60, 70]]
    assert md.cpu().tolist() == [[1.0, 2.0, 3.0], [5.0, 6.0, 7.0]]
    md, mi = b2.merge_topk(d, i, 2, descending=True)
    assert mi.cpu().tolist() == [[40, 30], [80, 70]]


def test_merge_topk_random_vs_oracle(b2):
    from oracle.merge import merge_topk
    g = torch.Generator().manual_seed(5)
    for parts, nq, k_in, k_out in [(8, 300, 100, 100), (3, 17, 7, 10), (2, 5, 200, 128), (5, 64, 1, 4)]:
        d = torch.sort(torch.rand(parts, nq, k_in, generator=g), dim=2).values
        i = torch.randint(0, 10**9, (parts, nq, k_in), generator=g)
        md, mi = b2.merge_topk(d.cuda(), i.cuda(), k_out)
        od, oi = merge_topk(list(d.numpy()), list(i.numpy()), k_out)
        kk = min(k_out, parts * k_in)
        np.testing.assert_array_equal(mi.cpu().numpy()[:, :kk], oi[:, :kk])
        np.testing.assert_array_equal(md.cpu().numpy()[:, :kk], od[:, :kk])


def test_sharded_search_equals_single_shard(b2):
    """Property at size: 1 shard vs 4 uneven shards + merge give identical ids (1M x 128)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    db = torch.randn(1_000_003, 128, generator=g, device="cuda").to(torch.bfloat16)
    qs = torch.randn(2048, 128, generator=g, device="cuda").to(torch.bfloat16)
    d1, i1 = b2.NativeIndex.flat(db).search(qs, 100)
    dd, ii = [], []
    for s, e in b2.partition_even(db.shape[0], 4):
        d, i = b2.NativeIndex.flat(db[s:e].contiguous(), id_offset=s).search(qs, 100)
        dd.append(d); ii.append(i)
    md, mi = b2.merge_topk(torch.stack(dd), torch.stack(ii), 100)
    torch.cuda.synchronize()
    same = (mi == i1).float().mean().item()
    assert same > 0.9995, same
    assert torch.allclose(md, d1, rtol=1e-4, atol=1e-3)
    assert (md[:, 1:] >= md[:, :-1]).all()                      # sortedness
    assert int((mi < 0).sum()) == 0 and int(mi.max()) < db.shape[0]
    # split invariance: forcing a different db decomposition must not change the answer
    d3, i3 = b2.NativeIndex.flat(db).search(qs, 100, n_splits=3)
    assert (i3 == i1).float().mean().item() > 0.9995


def test_three_pass_threshold_seeding_at_large_n(b2):
    """N >= 8192 tiles triggers the stride-256 / stride-16 / full pass chain; ids must still match
    an exact fp32 search (checked with the blocked CPU oracle on a query subset)."""
    from oracle.exact import exact_knn
    g = torch.Generator(device="cuda").manual_seed(17)
    n, d = 2_200_000, 64
    db = torch.randn(n, d, generator=g, device="cuda").to(torch.bfloat16)
    qs = torch.randn(384, d, generator=g, device="cuda").to(torch.bfloat16)
    ix = b2.NativeIndex.flat(db)
    for k in (10, 100):
        dd, ii = ix.search(qs, k)
        torch.cuda.synchronize()
        rd, ri = exact_knn(db.float().cpu(), qs[:48].float().cpu(), k)
        same = (ii[:48].cpu() == ri).float().mean().item()
        assert same > 0.999, (k, same)
        assert torch.allclose(dd[:48].cpu(), rd, rtol=1e-3, atol=1e-2)
        assert (dd[:, 1:] >= dd[:, :-1]).all()
        assert ix.last_stats().launches >= 6      # 3 fused launches + 3 merges (+ query norms)


def test_adversarial_order_sorted_database(b2):
    """Rows sorted by distance to the query direction (best rows LAST) defeat the sampled
    thresholds; the compaction fallback must keep the result exact."""
    from oracle.exact import topk_parity_report
    g = torch.Generator().manual_seed(23)
    n, d = 70000, 32
    db = torch.randn(n, d, generator=g)
    q = torch.randn(4, d, generator=g)
    order = torch.argsort(((db - q[0]) ** 2).sum(1), descending=True)
    db = db[order].to(torch.bfloat16)
    qs = q.to(torch.bfloat16)
    ix = b2.NativeIndex.flat(db.cuda())
    dd, ii = ix.search(qs.cuda(), 100)
    torch.cuda.synchronize()
    assert topk_parity_report(dd.cpu(), ii.cpu(), db.float(), qs.float(), 100, rtol=RTOL)["ok"]


# ---- two-pass selection (small database, large k: the IVF coarse-probe shape)
TWO_PASS_CASES = [
    # n, d, q, k, dtype, metric
    (16384, 128, 3000, 64, "fp16", "sqeuclidean"),   # C4 coarse probe
    (4096, 768, 1000, 32, "fp16", "sqeuclidean"),    # C3 coarse probe
    (5003, 96, 700, 50, "bf16", "inner_product"),    # ragged rows and queries
    (1000, 64, 129, 128, "fp32", "sqeuclidean"),     # fp32 source (three-term operand), k = maximum
    (300, 40, 5, 2, "bf16", "sqeuclidean"),          # a handful of queries, dim % 8 != 0
]


@pytest.mark.parametrize("n,d,q,k,dtype,metric", TWO_PASS_CASES)
def test_two_pass_selection_parity(b2, setenv, n, d, q, k, dtype, metric):
    """B2VS_TWO_PASS=1 routes the search through chunk minima -> exact threshold -> survivors:
    same answer as the oracle and as the fused path, with a sub-batch size that forces several
    query batches."""
    db, qs = make(n, d, q, dtype, metric, seed=77)
    setenv("B2VS_TWO_PASS", "1")
    setenv("B2VS_TWO_PASS_CHUNK_MB", "8")
    d1, i1, ix = check(b2, db, qs, k, metric, id_offset=500)
    assert ix.last_stats().launches >= 5
    setenv("B2VS_TWO_PASS", "0")
    d0, i0 = ix.search(qs.cuda(), k)
    torch.cuda.synchronize()
    same = (i0.cpu() == i1).float().mean().item()
    assert same > 0.999, same          # accumulation order is identical: only exact ties may differ
    assert torch.allclose(d0.cpu(), d1, rtol=1e-5, atol=1e-5)


def test_two_pass_default_heuristic_and_ties(b2, setenv):
    """The coarse-probe shape takes the two-pass path by default; duplicate rows (exact score
    ties, also at the threshold) come back smaller id first, and padding rows never appear."""
    g = torch.Generator().manual_seed(5)
    base = torch.randn(2048, 64, generator=g)
    db = torch.cat([base, base]).to(torch.float16)          # row i and row i + 2048 are identical
    qs = torch.randn(600, 64, generator=g).to(torch.float16)
    ix = b2.NativeIndex.flat(db.cuda())
    d1, i1 = ix.search(qs.cuda(), 32)
    torch.cuda.synchronize()
    i1 = i1.cpu()
    assert (i1 >= 0).all() and (i1 < 4096).all()
    assert (i1[:, 0::2] + 2048 == i1[:, 1::2]).all()        # every hit is followed by its duplicate
    assert (d1.cpu()[:, 0::2] == d1.cpu()[:, 1::2]).all()
    setenv("B2VS_TWO_PASS", "0")
    d0, i0 = ix.search(qs.cuda(), 32)
    torch.cuda.synchronize()
    assert (i0.cpu() == i1).all()


def test_two_pass_all_rows_identical(b2):
    """Degenerate ties: every row equal, so every chunk minimum ties at the threshold - the
    (score, chunk) threshold still admits at most 32 k rows and the answer is ids 0..k-1."""
    db = torch.ones(4096, 64, dtype=torch.float16) * 0.5
    qs = torch.randn(640, 64, generator=torch.Generator().manual_seed(3)).to(torch.float16)
    ix = b2.NativeIndex.flat(db.cuda())
    d1, i1 = ix.search(qs.cuda(), 48)
    torch.cuda.synchronize()
    assert (i1.cpu() == torch.arange(48).expand(640, 48)).all()


# ---- large k (the reference's top-2000 retrieval mode)
@pytest.mark.parametrize("k,metric", [(1000, "sqeuclidean"), (2000, "sqeuclidean"), (500, "inner_product")])
def test_large_k_exact(b2, k, metric):
    from oracle.exact import exact_knn
    db, qs = make(120_000, 64, 150, "bf16", metric)
    ix = b2.NativeIndex.flat(db.cuda(), metric=metric, id_offset=9)
    d, i = ix.search(qs.cuda(), k)
    torch.cuda.synchronize()
    from oracle.exact import topk_parity_report
    rep = topk_parity_report(d.cpu(), i.cpu() - 9, db.float(), qs.float(), k, metric, rtol=RTOL)
    assert rep["ok"], rep                       # ids exact up to ties within 1e-3 relative
    rd, ri = exact_knn(db.float(), qs.float(), k, metric)
    same_set = np.mean([len(set(a) & set(b)) / float(k) for a, b in zip((i.cpu() - 9).tolist(), ri.tolist())])
    assert same_set > 0.9995, same_set          # the neighbour SETS agree; only tie order may differ
    assert torch.allclose(d.cpu(), rd, rtol=1e-3, atol=1e-2)


def test_large_k_overflow_recovery(b2):
    """Sampled tiles hold only far rows, so the seeded threshold lets >65536 candidates through:
    the append buffers overflow and the pass must be repeated with the tightened threshold."""
    from oracle.exact import exact_knn
    g = torch.Generator().manual_seed(31)
    n, d, k = 400_000, 32, 1000
    db = torch.randn(n, d, generator=g)
    tile = torch.arange(n) // 256
    stride = -(-((n + 255) // 256 * 256) // 32768)           # first-pass tile stride used by the engine
    db[tile % stride == 0] += 40.0                           # the sample sees only far rows
    db = db.to(torch.bfloat16)
    qs = torch.randn(40, d, generator=g).to(torch.bfloat16)
    ix = b2.NativeIndex.flat(db.cuda())
    dd, ii = ix.search(qs.cuda(), k)
    torch.cuda.synchronize()
    rd, ri = exact_knn(db.float(), qs.float(), k)
    same_set = np.mean([len(set(a) & set(b)) / float(k) for a, b in zip(ii.cpu().tolist(), ri.tolist())])
    assert same_set > 0.9995, same_set
    assert torch.allclose(dd.cpu(), rd, rtol=1e-3, atol=1e-2)


def test_merge_topk_large_k(b2):
    from oracle.merge import merge_topk
    g = torch.Generator().manual_seed(6)
    d = torch.sort(torch.rand(4, 50, 2000, generator=g), dim=2).values
    i = torch.randint(0, 10**9, (4, 50, 2000), generator=g)
    md, mi = b2.merge_topk(d.cuda(), i.cuda(), 2000)
    od, oi = merge_topk(list(d.numpy()), list(i.numpy()), 2000)
    np.testing.assert_array_equal(mi.cpu().numpy(), oi)
    np.testing.assert_array_equal(md.cpu().numpy(), od)
