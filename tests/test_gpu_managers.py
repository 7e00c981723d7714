"""The drop-in pipeline on real GPUs: GRM -> EDM -> IBC -> SRA, results against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pipeline(b2, index_type, params, metric="sqeuclidean", dtype=torch.bfloat16, k=10, search_params=None):
    from oracle.exact import exact_knn
    g = torch.Generator().manual_seed(21)
    emb = torch.randn(30001, 64, generator=g)
    qs = torch.randn(50, 64, generator=g)
    grm = b2.GPUResourceManager()
    assert grm.get_available_gpu_count() >= 1
    edm = b2.EmbeddingDistributionManager(grm)
    dist = edm.distribute_embeddings(emb, dtype=dtype)
    assert edm.validate_distribution(dist)
    ibc = b2.IndexBuildingCoordinator(grm)
    cfg = b2.IndexBuildConfig(index_type, dict(params, metric=metric), max_retries=0)
    built = ibc.build_indices_parallel(dist, cfg)
    assert built.success, [r.error_message for r in built.build_results]
    sra = b2.SearchResultAggregator(grm)
    res = sra.perform_distributed_search(qs, ibc.get_built_indices(),
                                         b2.SearchConfig(k=k, search_params=search_params))
    rd, ri = exact_knn(emb.to(dtype).float(), qs.to(dtype).float(), k, metric)
    return res, rd, ri, ibc, edm


def test_brute_force_pipeline_is_exact(b2):
    res, rd, ri, ibc, edm = pipeline(b2, "brute_force", {})
    assert res.final_indices.shape == (50, 10) and res.num_queries == 50
    assert (res.final_indices == ri.numpy()).mean() > 0.995
    np.testing.assert_allclose(res.final_distances, rd.numpy(), rtol=1e-3, atol=1e-2)
    assert len(res.gpu_results) == len(ibc.get_built_indices())
    for r in res.gpu_results:
        assert r.distances.shape == (50, 10) and r.k_returned == 10
    ibc.cleanup_all_indices()
    edm.cleanup_distribution()


def test_inner_product_pipeline_descending(b2):
    res, rd, ri, ibc, _ = pipeline(b2, "flat", {}, metric="inner_product")
    assert (np.diff(res.final_distances, axis=1) <= 1e-6).all()
    assert (res.final_indices == ri.numpy()).mean() > 0.995


def test_ivf_flat_pipeline_default_nlists_and_probes(b2):
    res, rd, ri, ibc, _ = pipeline(b2, "ivf_flat", {}, search_params={"n_probes": 31})
    # default n_lists = min(256, N // 1000 + 1) = 31 per full corpus shard -> full probe = exact
    hit = np.mean([len(set(a) & set(b)) / 10.0 for a, b in zip(res.final_indices.tolist(), ri.tolist())])
    assert hit > 0.97, hit


def test_ivf_pq_pipeline_runs(b2):
    res, rd, ri, ibc, _ = pipeline(b2, "ivf_pq", {"n_lists": 16, "pq_dim": 32}, search_params={"nprobe": 16})
    hit = np.mean([len(set(a) & set(b)) / 10.0 for a, b in zip(res.final_indices.tolist(), ri.tolist())])
    assert hit > 0.5, hit


def test_cagra_is_reported_not_faked(b2):
    grm = b2.GPUResourceManager()
    ibc = b2.IndexBuildingCoordinator(grm)
    part = b2.EmbeddingPart(0, torch.randn(100, 16).cuda(), 0, 100)
    dist = b2.DistributedEmbeddings([part], 100, 16)
    out = ibc.build_indices_parallel(dist, b2.IndexBuildConfig("cagra", {}, max_retries=0))
    assert not out.success and "out of scope" in out.build_results[0].error_message


def test_simulated_index_is_rejected_on_a_gpu_box(b2):
    grm = b2.GPUResourceManager()
    sra = b2.SearchResultAggregator(grm)
    with pytest.raises(TypeError, match="not a native index"):
        sra.perform_distributed_search(torch.randn(2, 8), {0: {"type": "ivf_flat", "size": 1, "dim": 8}},
                                       b2.SearchConfig(k=1, parallel_search=False))


def test_device_to_device_reshard_reassembles_rows(b2):
    """Elastic re-shard without the host round trip: new shards are filled from the overlapping
    row ranges of the old ones by device-to-device copies."""
    grm = b2.GPUResourceManager()
    edm = b2.EmbeddingDistributionManager(grm)
    x = torch.randn(1000, 48)
    parts = [b2.EmbeddingPart(0, x[:600].cuda(), 0, 600), b2.EmbeddingPart(0, x[600:].cuda(), 600, 1000)]
    dist = b2.DistributedEmbeddings(parts, 1000, 48)
    out = edm._reshard_device_to_device(dist, [0])
    assert out is not None and len(out.parts) == 1
    p = out.parts[0]
    assert (p.gpu_id, p.start_index, p.end_index) == (0, 0, 1000) and p.tensor.is_cuda
    assert torch.equal(p.tensor.cpu(), x)
    assert parts[0].tensor is None and parts[1].tensor is None      # old shards were released
    assert edm.current_distribution is out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NVLink peer copies)")
def test_redistribute_moves_shards_over_nvlink_when_a_gpu_drops_out(b2):
    grm = b2.GPUResourceManager()
    edm = b2.EmbeddingDistributionManager(grm)
    x = torch.randn(5001, 64)
    dist = edm.distribute_embeddings(x, target_gpus=[0, 1], dtype=torch.bfloat16)
    assert [p.gpu_id for p in dist.parts] == [0, 1]
    grm.available_gpus = [1]                                         # GPU 0 is taken away
    out = edm.redistribute_if_needed(dist)
    assert [(p.gpu_id, p.start_index, p.end_index) for p in out.parts] == [(1, 0, 5001)]
    assert torch.equal(out.parts[0].tensor.cpu(), x.to(torch.bfloat16))
    # and the index built on the re-sharded data answers with global ids
    ix = b2.NativeIndex.flat(out.parts[0].tensor, metric="sqeuclidean")
    _, ids = ix.search(out.parts[0].tensor[[7, 4999]], 1)
    assert ids[:, 0].tolist() == [7, 4999]


def test_staged_upload_equals_plain_copy(b2, monkeypatch):
    """EDM shards of 64 MB and more go up through two pinned staging buffers in chunks, converted
    on the device chunk by chunk; the resident shard must equal the one-shot copy bit for bit."""
    edm_mod = b2.embedding_distribution_manager
    g = torch.Generator().manual_seed(3)
    x = torch.randn(50_001, 96, generator=g)                       # 19.2 MB, ragged last chunk
    dev = torch.device("cuda:0")
    for dtype in (None, torch.bfloat16, torch.float16):
        got = edm_mod.staged_host_to_device(x, dev, dtype, chunk_bytes=1 << 20)     # 19 chunks
        want = x.to(dev) if dtype is None else x.to(dev).to(dtype)
        assert got.dtype == want.dtype and got.is_contiguous() and torch.equal(got, want)
    # pinned sources skip the staging buffers
    got = edm_mod.staged_host_to_device(x.pin_memory(), dev, torch.bfloat16, chunk_bytes=1 << 20)
    assert torch.equal(got, x.to(dev).to(torch.bfloat16))
    # through the manager (threshold lowered so this small matrix takes the staged route)
    monkeypatch.setattr(edm_mod, "STAGE_MIN_BYTES", 1)
    monkeypatch.setattr(edm_mod, "STAGE_CHUNK_BYTES", 1 << 20)
    grm = b2.GPUResourceManager()
    edm = b2.EmbeddingDistributionManager(grm)
    dist = edm.distribute_embeddings(x, dtype=torch.bfloat16)
    assert edm.validate_distribution(dist)
    for part in dist.parts:
        want = x[part.start_index:part.end_index].to(part.tensor.device).to(torch.bfloat16)
        assert torch.equal(part.tensor, want)
    edm.cleanup_distribution()
