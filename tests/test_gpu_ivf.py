"""GPU IVF-Flat / IVF-PQ / k-means through the C ABI.  Parity criterion (north star): recall@k at
the same n_lists / n_probes must match the restated reference-semantics index (oracle.ivf) within
sampling noise; a full probe of an IVF-Flat index must reproduce the exact search."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def clustered(n, d, c, seed, sigma=0.6):
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(c, d, generator=g)
    return cent[torch.randint(0, c, (n,), generator=g)] + sigma * torch.randn(n, d, generator=g)


def queries_from(x, nq, seed):
    g = torch.Generator().manual_seed(seed)
    return x[torch.randperm(x.shape[0], generator=g)[:nq]] + 0.1 * torch.randn(nq, x.shape[1], generator=g)


def test_kmeans_assignment_and_inertia(b2):
    x = clustered(50000, 64, 32, 1, sigma=0.2)
    for dt in (torch.float16, torch.bfloat16, torch.float32):
        xg = x.to(dt).cuda()
        c, lab = b2.kmeans_fit(xg, 32, iters=15, seed=3)
        torch.cuda.synchronize()
        xf, cf = xg.float(), (c if dt == torch.float32 else c.to(dt).float())
        dist = (cf * cf).sum(1)[None, :] - 2.0 * xf @ cf.T
        chosen = dist.gather(1, lab.long()[:, None])[:, 0]
        # every row sits on (numerically) its nearest centroid
        assert float((chosen - dist.min(1).values).max()) < 1e-2
        inertia = ((xf - c[lab.long()]) ** 2).sum(1).mean().item()
        assert inertia < 64 * 0.04 * 6, inertia     # Lloyd local optimum, far below the 66 of one blob


@pytest.mark.parametrize("dtype,metric,d", [(torch.float16, "sqeuclidean", 128),
                                             (torch.bfloat16, "inner_product", 128),
                                             (torch.float32, "sqeuclidean", 96)])
def test_ivf_flat_recall_matches_oracle(b2, dtype, metric, d):
    from oracle.exact import exact_knn
    from oracle.ivf import IvfFlatOracle, recall
    n, nlist, nprobe, k = 60000, 128, 8, 10
    x = clustered(n, d, 150, 5).to(dtype)
    q = queries_from(x.float(), 300, 9).to(dtype)
    ix = b2.NativeIndex.ivf_flat(x.cuda(), nlist, metric=metric, id_offset=11, kmeans_iters=10)
    sizes = ix.list_sizes()
    assert int(sizes.sum()) == n and ix.info().n_lists == nlist
    _, ti = exact_knn(x.float(), q.float(), k, metric)
    _, gi = ix.search(q.cuda(), k, n_probes=nprobe)
    r_gpu = recall(gi.cpu() - 11, ti)
    oracle = IvfFlatOracle(x.float(), nlist, metric, iters=10)
    _, oi = oracle.search(q.float(), k, n_probes=nprobe)
    r_ref = recall(oi, ti)
    assert abs(r_gpu - r_ref) < 0.05, (r_gpu, r_ref)
    assert r_gpu > 0.8
    # probing every list is an exact search
    _, fi = ix.search(q.cuda(), k, n_probes=nlist)
    assert recall(fi.cpu() - 11, ti) > 0.995


def test_ivf_flat_distances_are_exact_for_returned_ids(b2):
    x = clustered(20000, 64, 50, 2).to(torch.float16)
    q = queries_from(x.float(), 50, 3).to(torch.float16)
    ix = b2.NativeIndex.ivf_flat(x.cuda(), 64, kmeans_iters=5)
    d, i = ix.search(q.cuda(), 5, n_probes=4)
    d, i = d.cpu(), i.cpu()
    true = ((x.float()[i.clamp_min(0)] - q.float()[:, None, :]) ** 2).sum(2)
    ok = i >= 0
    assert torch.allclose(d[ok], true[ok], rtol=2e-3, atol=2e-3)
    assert (d[:, 1:] >= d[:, :-1] - 1e-4).all()


@pytest.mark.parametrize("metric", ["sqeuclidean", "inner_product"])
def test_ivf_pq_recall_matches_oracle(b2, metric):
    from oracle.exact import exact_knn
    from oracle.ivf import IvfPqOracle, recall
    n, d, nlist, nprobe, k, m = 40000, 64, 64, 16, 10, 32
    x = clustered(n, d, 100, 6).to(torch.float16)
    q = queries_from(x.float(), 200, 7).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, m, metric=metric, kmeans_iters=10)
    inf = ix.info()
    assert (inf.pq_dim, inf.pq_bits, inf.n_lists) == (m, 8, nlist)
    _, ti = exact_knn(x.float(), q.float(), k, metric)
    _, gi = ix.search(q.cuda(), k, n_probes=nprobe)
    r_gpu = recall(gi.cpu(), ti)
    oracle = IvfPqOracle(x.float(), nlist, m, metric, iters=10, pq_iters=10)
    _, oi = oracle.search(q.float(), k, n_probes=nprobe)
    r_ref = recall(oi, ti)
    assert abs(r_gpu - r_ref) < 0.06, (r_gpu, r_ref)
    assert r_gpu > 0.6


def test_ivf_build_argument_errors(b2):
    x = torch.randn(1000, 32).half().cuda()
    with pytest.raises(RuntimeError, match="n_lists"):
        b2.NativeIndex.ivf_flat(x, 5000)
    with pytest.raises(RuntimeError, match="must divide dim"):
        b2.NativeIndex.ivf_pq(x, 8, 5)


def test_ivf_pq_refine_lifts_recall_like_the_oracle(b2):
    from oracle.exact import exact_knn
    from oracle.ivf import IvfPqOracle, recall
    n, d, nlist, nprobe, k, m = 40000, 64, 64, 16, 10, 16     # dsub = 4: coarse codes
    x = clustered(n, d, 100, 8).to(torch.float16)
    q = queries_from(x.float(), 200, 9).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, m, kmeans_iters=10, id_offset=3)
    _, ti = exact_knn(x.float(), q.float(), k)
    _, g0 = ix.search(q.cuda(), k, n_probes=nprobe)
    d1, g1 = ix.search(q.cuda(), k, n_probes=nprobe, refine_ratio=8)
    r0, r1 = recall(g0.cpu() - 3, ti), recall(g1.cpu() - 3, ti)
    oracle = IvfPqOracle(x.float(), nlist, m, iters=10, pq_iters=10)
    _, o1 = oracle.search(q.float(), k, n_probes=nprobe, refine_ratio=8)
    ro = recall(o1, ti)
    assert r1 > r0 + 0.05 and abs(r1 - ro) < 0.06, (r0, r1, ro)
    # refined distances are exact for the returned ids
    d1, g1 = d1.cpu(), g1.cpu() - 3
    true = ((x.float()[g1.clamp_min(0)] - q.float()[:, None, :]) ** 2).sum(2)
    assert torch.allclose(d1, true, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("d,m,label", [(128, 64, "specialised query-major scan (M=64, dsub=2)"),
                                       (768, 96, "per-(query, probe) scan: codebooks exceed smem")])
def test_ivf_pq_kernel_variants_match_oracle(b2, setenv, d, m, label):
    from oracle.exact import exact_knn
    from oracle.ivf import IvfPqOracle, recall
    setenv("B2VS_IVF_GROUPED", "0")      # the look-up-table kernels, not the grouped scan
    n, nlist, nprobe, k = 20000, 32, 8, 10
    x = clustered(n, d, 60, 12).to(torch.float16)
    q = queries_from(x.float(), 100, 13).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, m, kmeans_iters=8)
    _, ti = exact_knn(x.float(), q.float(), k)
    _, gi = ix.search(q.cuda(), k, n_probes=nprobe, refine_ratio=4)
    oracle = IvfPqOracle(x.float(), nlist, m, iters=8, pq_iters=8)
    _, oi = oracle.search(q.float(), k, n_probes=nprobe, refine_ratio=4)
    r_gpu, r_ref = recall(gi.cpu(), ti), recall(oi, ti)
    assert abs(r_gpu - r_ref) < 0.07, (label, r_gpu, r_ref)


@pytest.mark.parametrize("kind", ["ivf_flat", "ivf_pq"])
def test_index_save_load_roundtrip(b2, tmp_path, kind):
    x = clustered(30000, 64, 60, 14).to(torch.float16).cuda()
    q = queries_from(x.float().cpu(), 64, 15).to(torch.float16).cuda()
    if kind == "ivf_flat":
        ix = b2.NativeIndex.ivf_flat(x, 48, id_offset=100, kmeans_iters=6)
        kw = dict(n_probes=8)
    else:
        ix = b2.NativeIndex.ivf_pq(x, 48, 32, id_offset=100, kmeans_iters=6)
        kw = dict(n_probes=8, refine_ratio=4)
    d0, i0 = ix.search(q, 10, **kw)
    path = str(tmp_path / f"{kind}.b2vs")
    ix.save(path)
    ix.destroy()
    ix2 = b2.NativeIndex.load(path, "cuda:0", rows=x if kind == "ivf_pq" else None)
    inf = ix2.info()
    assert (inf.n_rows, inf.dim, inf.n_lists, inf.id_offset) == (30000, 64, 48, 100)
    d1, i1 = ix2.search(q, 10, **kw)
    assert torch.equal(i0, i1) and torch.allclose(d0, d1)        # bit-identical after reload
    assert int(ix2.list_sizes().sum()) == 30000
    ix3 = b2.NativeIndex.load(path, "cuda:0", rows=x if kind == "ivf_pq" else None, id_offset=0)
    _, i3 = ix3.search(q, 10, **kw)
    assert torch.equal(i3, torch.where(i0 >= 0, i0 - 100, i0))


def test_flat_index_save_is_refused(b2, tmp_path):
    ix = b2.NativeIndex.flat(torch.randn(100, 16).half().cuda())
    with pytest.raises(RuntimeError, match="re-create"):
        ix.save(str(tmp_path / "flat.b2vs"))
    with pytest.raises(RuntimeError, match="not a b2vs index file"):
        p = tmp_path / "junk.b2vs"
        p.write_bytes(b"x" * 4096)
        b2.NativeIndex.load(str(p), "cuda:0")


@pytest.mark.parametrize("dtype,metric,d,k", [(torch.bfloat16, "sqeuclidean", 128, 10),
                                               (torch.float16, "inner_product", 72, 10),
                                               (torch.bfloat16, "sqeuclidean", 200, 100),
                                               (torch.float32, "sqeuclidean", 200, 100),
                                               (torch.float32, "inner_product", 96, 10)])
def test_ivf_flat_grouped_scan_equals_per_item_scan(b2, setenv, dtype, metric, d, k):
    """Large batches take the grouped tensor-core list scan; it must return what the per-(query,
    probe) scan returns on the same index (ties aside), including when every candidate buffer
    overflows and the rescue kernel answers instead."""
    from oracle.exact import topk_parity_report
    n, nlist, nprobe, nq = 40000, 64, 12, 700
    x = clustered(n, d, 80, 21).to(dtype)
    q = queries_from(x.float(), nq, 22).to(dtype)
    ix = b2.NativeIndex.ivf_flat(x.cuda(), nlist, metric=metric, id_offset=5, kmeans_iters=8)
    setenv("B2VS_IVF_GROUPED", "0")
    d0, i0 = ix.search(q.cuda(), k, n_probes=nprobe)
    torch.cuda.synchronize()
    setenv("B2VS_IVF_GROUPED", "1")
    d1, i1 = ix.search(q.cuda(), k, n_probes=nprobe)
    torch.cuda.synchronize()
    setenv("B2VS_IVF_GROUPED_CAP", "32")     # forces overflow -> rescue path
    d2, i2 = ix.search(q.cuda(), k, n_probes=nprobe)
    torch.cuda.synchronize()
    for dd, ii in ((d1, i1), (d2, i2)):
        scale = float(d0.abs().max())
        assert float((dd - d0).abs().max()) <= 2e-3 * scale + 1e-3
        same = (ii == i0).float().mean().item()
        assert same > 0.995, same
        # rows that differ are ties / fp-order swaps: same id sets almost everywhere
        inter = [len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ii.cpu(), i0.cpu())]
        assert sum(inter) >= 0.999 * nq * k


def test_ivf_flat_grouped_scan_ragged_batch(b2, setenv):
    """Batch sizes that are not multiples of the 128-row query block and lists nobody probes."""
    from oracle.exact import exact_knn
    from oracle.ivf import recall
    x = clustered(20000, 64, 40, 31).to(torch.bfloat16)
    ix = b2.NativeIndex.ivf_flat(x.cuda(), 32, metric="sqeuclidean", kmeans_iters=8)
    setenv("B2VS_IVF_GROUPED", "1")
    for nq in (1, 65, 129):
        q = queries_from(x.float(), nq, 40 + nq).to(torch.bfloat16)
        _, ti = exact_knn(x.float(), q.float(), 5, "sqeuclidean")
        dd, ii = ix.search(q.cuda(), 5, n_probes=32)
        assert recall(ii.cpu(), ti) > 0.99
        assert bool((dd[:, 1:] >= dd[:, :-1]).all())


@pytest.mark.parametrize("metric,d,m", [("sqeuclidean", 128, 64), ("inner_product", 128, 64),
                                         ("sqeuclidean", 64, 16),      # dsub 4, codebooks in smem
                                         ("sqeuclidean", 128, 16),     # dsub 8, codebooks in smem
                                         ("sqeuclidean", 256, 32),     # dsub 8, codebooks via L2
                                         ("inner_product", 256, 128)]) # dsub 2, codebooks via L2
def test_ivf_pq_grouped_scan_equals_lut_scan(b2, setenv, metric, d, m):
    """Large batches decode each probed list once into bf16 tiles for the tensor cores.  Its ADC
    scores use bf16-rounded residual queries / codebooks, so against the fp32 look-up-table scan
    of the same index: near-identical candidate sets and distances, same recall; after the exact
    refine step the answers coincide.  Also drives the overflow-rescue kernel."""
    from oracle.exact import exact_knn
    from oracle.ivf import recall
    n, nlist, nprobe, nq, k = 40000, 64, 12, 600, 10
    x = clustered(n, d, 80, 51).to(torch.float16)
    q = queries_from(x.float(), nq, 52).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, m, metric=metric, kmeans_iters=8, id_offset=7)
    _, ti = exact_knn(x.float(), q.float(), k, metric)
    out = {}
    for mode in ("0", "1"):
        setenv("B2VS_IVF_GROUPED", mode)
        out[mode] = ix.search(q.cuda(), k, n_probes=nprobe)
        out[mode + "r"] = ix.search(q.cuda(), k, n_probes=nprobe, refine_ratio=4)
        torch.cuda.synchronize()
    setenv("B2VS_IVF_GROUPED_CAP", "32")
    out["1c"] = ix.search(q.cuda(), k, n_probes=nprobe)
    torch.cuda.synchronize()
    (d0, i0), (d1, i1), (d2, i2) = out["0"], out["1"], out["1c"]
    for dd, ii in ((d1, i1), (d2, i2)):
        inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ii.cpu(), i0.cpu()))
        assert inter >= 0.93 * nq * k, inter / (nq * k)
        scale = float(d0[torch.isfinite(d0)].abs().max())
        assert float((dd - d0).abs().median()) <= 5e-3 * scale
        assert abs(recall(ii.cpu() - 7, ti) - recall(i0.cpu() - 7, ti)) < 0.02
    # the rescue kernel uses the same operands as the tensor-core path: same answer as uncapped
    assert (i2 == i1).float().mean().item() > 0.97
    (dr0, ir0), (dr1, ir1) = out["0r"], out["1r"]
    assert abs(recall(ir1.cpu() - 7, ti) - recall(ir0.cpu() - 7, ti)) < 0.01
    assert (ir1 == ir0).float().mean().item() > 0.97


def test_ivf_search_runs_huge_batches_in_sub_batches(b2):
    """nq * n_probes beyond the workspace budget is processed as consecutive sub-batches."""
    from oracle.ivf import recall
    x = clustered(20000, 32, 50, 61).to(torch.bfloat16).cuda()
    ix = b2.NativeIndex.ivf_flat(x, 128, metric="sqeuclidean", kmeans_iters=5)
    nq = 40000                                   # 128 probes -> sub-batches of 32768 queries
    q = x[torch.arange(nq, device="cuda") % x.shape[0]].clone()
    dd, ii = ix.search(q, 3, n_probes=128)
    flat = b2.NativeIndex.flat(x, metric="sqeuclidean")
    fd, fi = flat.search(q, 3)
    assert recall(ii.cpu(), fi.cpu()) > 0.999
    assert torch.allclose(dd, fd, rtol=1e-3, atol=1e-2)
    assert float(dd[:, 0].max()) < 1e-2          # every query is a database row


@pytest.mark.parametrize("nq", [40, 300])
def test_ivf_more_than_128_probes(b2, nq):
    """n_probes above the fused top-k limit (128) goes through the large-k coarse probe; probing
    every one of 256 lists must reproduce the exact search (both scan paths, Flat and PQ run)."""
    from oracle.ivf import recall
    x = clustered(30000, 64, 70, 71).to(torch.float16).cuda()
    q = queries_from(x.float().cpu(), nq, 72).to(torch.float16).cuda()
    flat = b2.NativeIndex.flat(x, metric="sqeuclidean")
    fd, fi = flat.search(q, 10)
    ix = b2.NativeIndex.ivf_flat(x, 256, metric="sqeuclidean", kmeans_iters=6)
    dd, ii = ix.search(q, 10, n_probes=256)
    assert recall(ii.cpu(), fi.cpu()) > 0.999
    assert torch.allclose(dd, fd, rtol=1e-3, atol=1e-2)
    d2, i2 = ix.search(q, 10, n_probes=200)
    assert recall(i2.cpu(), fi.cpu()) > 0.99
    pq = b2.NativeIndex.ivf_pq(x, 256, 32, metric="sqeuclidean", kmeans_iters=6)
    _, ip = pq.search(q, 10, n_probes=256, refine_ratio=8)
    assert recall(ip.cpu(), fi.cpu()) > 0.9


@pytest.mark.parametrize("kind", ["flat", "pq"])
def test_ivf_grouped_scan_with_every_query_on_the_same_lists(b2, setenv, kind):
    """Worst-case skew: 700 near-identical queries probe the same few lists, so those lists get
    several 128-row query blocks each and every other list none."""
    x = clustered(30000, 64, 60, 81).to(torch.bfloat16)
    q = (x[123].float()[None, :] + 0.01 * torch.randn(700, 64)).to(torch.bfloat16)
    if kind == "flat":
        ix = b2.NativeIndex.ivf_flat(x.cuda(), 64, metric="sqeuclidean", kmeans_iters=6)
    else:
        ix = b2.NativeIndex.ivf_pq(x.cuda(), 64, 32, metric="sqeuclidean", kmeans_iters=6)
    res = {}
    for mode in ("0", "1"):
        setenv("B2VS_IVF_GROUPED", mode)
        res[mode] = ix.search(q.cuda(), 10, n_probes=6)
        torch.cuda.synchronize()
    (d0, i0), (d1, i1) = res["0"], res["1"]
    inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i1.cpu(), i0.cpu()))
    assert inter >= (0.999 if kind == "flat" else 0.93) * 700 * 10
    assert bool((i1[:, 0] == 123).float().mean() > 0.95)          # the perturbed row itself
    assert bool((d1[:, 1:] >= d1[:, :-1]).all())


@pytest.mark.parametrize("dtype,metric,k", [(torch.bfloat16, "sqeuclidean", 500),
                                             (torch.float32, "inner_product", 2000)])
def test_ivf_flat_large_k(b2, dtype, metric, k):
    """128 < k <= 2048 on an IVF-Flat index (the reference's top-2000 mode): probing every list
    must reproduce the flat index's large-k answer; a partial probe is a subset search."""
    from oracle.ivf import recall
    x = clustered(40000, 96, 50, 91).to(dtype).cuda()
    q = queries_from(x.float().cpu(), 130, 92).to(dtype).cuda()
    ix = b2.NativeIndex.ivf_flat(x, 32, metric=metric, id_offset=9, kmeans_iters=6)
    dd, ii = ix.search(q, k, n_probes=32)
    torch.cuda.synchronize()
    assert ii.shape == (130, k) and int(ii.min()) >= 9
    # exact ground truth over the index's own representation (bf16 rows, fp32 accumulate)
    xf = x.to(torch.bfloat16).float() if dtype == torch.float32 else x.float()
    qf = q.float()
    if metric == "sqeuclidean":
        full = (xf * xf).sum(1)[None, :] - 2.0 * qf @ xf.T + (qf * qf).sum(1)[:, None]
        td, ti = torch.topk(full, k, dim=1, largest=False)
    else:
        full = qf @ xf.T
        td, ti = torch.topk(full, k, dim=1, largest=True)
    assert recall((ii - 9).cpu(), ti.cpu()) > 0.998
    assert torch.allclose(dd, td, rtol=2e-3, atol=2e-2)
    srt = dd[:, 1:] >= dd[:, :-1] if metric == "sqeuclidean" else dd[:, 1:] <= dd[:, :-1]
    assert bool(srt.all())
    d2, i2 = ix.search(q, k, n_probes=8)          # subset of the lists: still k results, sorted
    assert int((i2 >= 0).sum()) > 0.9 * i2.numel()
    pq = b2.NativeIndex.ivf_pq(x, 32, 24, metric=metric)
    with pytest.raises(RuntimeError, match="needs the grouped scan"):
        pq.search(q, k)                               # dim 96 is not a multiple of 64


@pytest.mark.parametrize("metric", ["sqeuclidean", "inner_product"])
def test_ivf_pq_large_k_and_deep_refine(b2, metric):
    """IVF-PQ beyond the fused limit: k = 500 ADC results, and k = 100 refined from 400 candidates
    (k * refine_ratio > 128 takes the large-k path instead of being clamped to 128)."""
    from oracle.exact import exact_knn
    from oracle.ivf import recall
    x = clustered(40000, 128, 50, 93).to(torch.float16)
    q = queries_from(x.float(), 150, 94).to(torch.float16)
    ix = b2.NativeIndex.ivf_pq(x.cuda(), 32, 64, metric=metric, id_offset=4, kmeans_iters=8)
    _, t500 = exact_knn(x.float(), q.float(), 500, metric)
    dd, ii = ix.search(q.cuda(), 500, n_probes=32)
    torch.cuda.synchronize()
    assert ii.shape == (150, 500) and int(ii.min()) >= 4
    srt = dd[:, 1:] >= dd[:, :-1] if metric == "sqeuclidean" else dd[:, 1:] <= dd[:, :-1]
    assert bool(srt.all())
    assert recall((ii - 4).cpu(), t500) > 0.8             # ADC ordering of the true top-500
    d1, i1 = ix.search(q.cuda(), 100, n_probes=32, refine_ratio=4)    # 400 candidates re-ranked
    d0, i0 = ix.search(q.cuda(), 100, n_probes=32, refine_ratio=1)
    _, t100 = exact_knn(x.float(), q.float(), 100, metric)
    r1, r0 = recall((i1 - 4).cpu(), t100), recall((i0 - 4).cpu(), t100)
    assert r1 > 0.97 and r1 > r0 + 0.02, (r0, r1)
    xs, qs = x.float(), q.float()
    g = (i1 - 4).cpu().clamp_min(0)
    true = ((xs[g] - qs[:, None, :]) ** 2).sum(2) if metric == "sqeuclidean" else (xs[g] * qs[:, None, :]).sum(2)
    assert torch.allclose(d1.cpu(), true, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("kind", ["flat", "pq"])
def test_small_batch_cuda_graph_replay_equals_direct_search(b2, kind):
    """B2VS_FLAG_GRAPH: from the second call of a signature on the search is one captured CUDA
    graph replayed against staging buffers; answers must be those of the direct launches, for
    fresh query / output tensors on every call, and must survive a workspace re-allocation
    caused by a larger batch in between (stale-pointer guard)."""
    n, d, nlist = 40000, 64, 64
    x = clustered(n, d, 100, 21).to(torch.float16).cuda()
    if kind == "flat":
        ix = b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=5, id_offset=5)
        kw = dict(n_probes=8)
    else:
        ix = b2.NativeIndex.ivf_pq(x, nlist, 32, kmeans_iters=5, id_offset=5)
        kw = dict(n_probes=8, refine_ratio=4)
    for nq in (1, 3, 64):
        qs = [queries_from(x.float().cpu(), nq, 30 + t).to(torch.float16).cuda() for t in range(4)]
        want = [tuple(t.clone() for t in ix.search(q, 10, **kw)) for q in qs]
        got = [tuple(t.clone() for t in ix.search(q, 10, graph=True, **kw)) for q in qs]
        torch.cuda.synchronize()
        for (wd, wi), (gd, gi) in zip(want, got):
            assert torch.equal(wi, gi)
            assert torch.allclose(wd, gd, rtol=1e-6, atol=0)
        assert ix.last_stats().launches > 0   # nodes of the captured graph
    # a larger batch grows the workspaces; the cached graphs must notice and re-capture
    q1 = queries_from(x.float().cpu(), 1, 77).to(torch.float16).cuda()
    wd, wi = (t.clone() for t in ix.search(q1, 10, **kw))
    big = queries_from(x.float().cpu(), 3000, 78).to(torch.float16).cuda()
    ix.search(big, 10, **kw)
    for _ in range(3):
        gd, gi = ix.search(q1, 10, graph=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(wi, gi) and torch.allclose(wd, gd, rtol=1e-6, atol=0)
    # searches on a side stream replay the same graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sd, si = ix.search(q1, 10, graph=True, stream=s, **kw)
    s.synchronize()
    assert torch.equal(wi, si) and torch.allclose(wd, sd, rtol=1e-6, atol=0)


@pytest.mark.parametrize("kind", ["flat", "pq"])
def test_large_batch_graph_replay_and_plan_overlap_equal_the_in_line_search(b2, setenv, kind):
    """Batches of 2048 queries and more are replayed as CUDA graphs by default, with the main pass
    planned on a side stream into the second set of planning buffers while the seed pass runs.
    Direct launches (B2VS_GRAPH=0), with and without the overlap, must give the same answers as
    the default path on fresh queries in every call (first call direct, capture, replays)."""
    n, d, nlist = 120000, 128, 256
    x = clustered(n, d, 300, 31).to(torch.float16).cuda()
    if kind == "flat":
        ix = b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=5, id_offset=9)
        kw = dict(n_probes=12)
    else:
        ix = b2.NativeIndex.ivf_pq(x, nlist, 64, kmeans_iters=5, id_offset=9)
        kw = dict(n_probes=12, refine_ratio=2)
    qs = [queries_from(x.float().cpu(), 2500, 50 + t).to(torch.float16).cuda() for t in range(5)]
    setenv("B2VS_GRAPH", "0")
    setenv("B2VS_PLAN_OVERLAP", "0")
    want = [tuple(t.clone() for t in ix.search(q, 10, **kw)) for q in qs]
    setenv("B2VS_PLAN_OVERLAP", None)
    overlapped = [tuple(t.clone() for t in ix.search(q, 10, **kw)) for q in qs]
    setenv("B2VS_GRAPH", None)
    graphed = [tuple(t.clone() for t in ix.search(q, 10, **kw)) for q in qs]
    torch.cuda.synchronize()
    for (wd, wi), (od, oi), (gd, gi) in zip(want, overlapped, graphed):
        assert torch.equal(wi, oi) and torch.equal(wi, gi)
        assert torch.allclose(wd, od, rtol=1e-6, atol=0) and torch.allclose(wd, gd, rtol=1e-6, atol=0)
    # a different batch size in between (workspaces may move), then the first size again
    ix.search(queries_from(x.float().cpu(), 5000, 99).to(torch.float16).cuda(), 10, **kw)
    gd, gi = ix.search(qs[0], 10, **kw)
    torch.cuda.synchronize()
    assert torch.equal(want[0][1], gi)


@pytest.mark.parametrize("kind,dtype,metric,d", [("flat", torch.float16, "sqeuclidean", 128),
                                                  ("flat", torch.bfloat16, "inner_product", 72),
                                                  ("flat", torch.float32, "sqeuclidean", 200),
                                                  ("pq", torch.float16, "sqeuclidean", 64)])
def test_tiny_batch_coarse_probe_scan_equals_tensor_core_probe(b2, setenv, kind, dtype, metric, d):
    """Batches of up to 32 queries pick their probe lists with the CUDA-core scan over the centroid
    operand matrix (K4b) instead of the tensor-core probe; both must lead to the same answers
    (B2VS_COARSE_SCAN=0 forces the tensor-core probe), including with fewer centroids than one
    256-row chunk and with a ragged last chunk."""
    for nlist in (40, 300):
        x = clustered(30000, d, 120, 41).to(dtype)
        if kind == "flat":
            ix = b2.NativeIndex.ivf_flat(x.cuda(), nlist, metric=metric, id_offset=3, kmeans_iters=6)
            kw = dict(n_probes=min(nlist, 24))
        else:
            ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, 32, metric=metric, id_offset=3, kmeans_iters=6)
            kw = dict(n_probes=min(nlist, 24), refine_ratio=4)
        for nq in (1, 5, 8, 32):
            q = queries_from(x.float(), nq, 50 + nq).to(dtype).cuda()
            setenv("B2VS_COARSE_SCAN", "0")
            d0, i0 = (t.clone() for t in ix.search(q, 10, **kw))
            setenv("B2VS_COARSE_SCAN", None)
            d1, i1 = (t.clone() for t in ix.search(q, 10, **kw))
            torch.cuda.synchronize()
            # same probe lists -> same candidates; a near-tie between two centroids may swap the
            # last probed list, which can only change an answer's tail
            same = (i0 == i1).float().mean().item()
            assert same >= 0.9, (nlist, nq, same)
            assert torch.equal(i0[:, 0], i1[:, 0])
            scale = float(d0.abs().max())
            assert float((d0[:, 0] - d1[:, 0]).abs().max()) <= 1e-3 * scale + 1e-4
        # a full probe is exact whichever way the lists were ranked
        if kind == "flat" and nlist == 40:
            qf = queries_from(x.float(), 4, 77).to(dtype).cuda()
            setenv("B2VS_COARSE_SCAN", "0")
            _, j0 = ix.search(qf, 10, n_probes=nlist)
            setenv("B2VS_COARSE_SCAN", None)
            _, j1 = ix.search(qf, 10, n_probes=nlist)
            assert torch.equal(j0, j1)


@pytest.mark.parametrize("kind", ["ivf_flat", "ivf_pq"])
def test_grouped_scan_seed_modes_agree_on_structureless_data(b2, setenv, kind):
    """Thresholds of the grouped scans come from the CUDA-core seed kernels (B2VS_IVF_SEED=0, small
    batches) or from a tensor-core seed pass over the heads of the nearest lists (=1, default from
    256 queries): on iid Gaussian rows - where the nearest list's head is a weak bound - both must
    give what the per-(query, probe) scan gives, the second without flooding the buffers."""
    g = torch.Generator().manual_seed(61)
    n, d, nlist, nprobe, nq, k = 60000, 128, 128, 32, 300, 10
    x = torch.randn(n, d, generator=g).to(torch.float16)
    q = torch.randn(nq, d, generator=g).to(torch.float16)
    if kind == "ivf_flat":
        ix = b2.NativeIndex.ivf_flat(x.cuda(), nlist, kmeans_iters=6)
        kw = {}
    else:
        ix = b2.NativeIndex.ivf_pq(x.cuda(), nlist, 64, kmeans_iters=6)
        kw = {"refine_ratio": 4}
    setenv("B2VS_IVF_GROUPED", "0")
    _, i_ref = ix.search(q.cuda(), k, n_probes=nprobe, **kw)
    setenv("B2VS_IVF_GROUPED", "1")
    cands = {}
    for seed in ("0", "1"):
        setenv("B2VS_IVF_SEED", seed)
        _, ii = ix.search(q.cuda(), k, n_probes=nprobe, **kw)
        torch.cuda.synchronize()
        cands[seed] = ix.last_stats().mean_candidates
        inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ii.cpu(), i_ref.cpu()))
        assert inter >= (0.999 if kind == "ivf_flat" else 0.97) * nq * k, (seed, inter / (nq * k))
    # the larger, tensor-core sample gives the tighter threshold
    assert 0 < cands["1"] < cands["0"], cands
