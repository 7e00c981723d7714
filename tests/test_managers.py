"""Host-side behaviour of the drop-in manager classes (no GPU): same observable contract as the
reference's Attempt_1/test_*.py, plus the fixes SURVEY.md §3.6 / Appendix B call for."""
import json
import os
from unittest.mock import MagicMock, Mock, patch

import numpy as np
import pytest
import torch


@pytest.fixture
def mods(b2):
    return b2


def make_grm(b2, gpus):
    m = b2.GPUResourceManager()
    m.available_gpus = list(gpus)
    return m


# ---------------------------------------------------------------- GPUResourceManager
def test_partition_matches_reference_implementation(b2, golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "partition.json")))
    checked = 0
    for c in cases:
        if c["n"] < c["g"]:
            continue
        m = make_grm(b2, range(c["g"]))
        if c["strategy"] == "memory_based":
            m.gpu_memory_info = {i: {"available": v} for i, v in enumerate(c["mem"])}
        got = [list(t) for t in m.distribute_workload(c["n"], c["strategy"])]
        assert got == c["out"], c
        checked += 1
    assert checked > 40


def test_partition_subset_covers_whole_range(b2):
    m = make_grm(b2, [0, 1, 2, 3])
    out = m.distribute_workload(10, "even", gpu_ids=[1, 3])
    assert out == [(1, 0, 5), (3, 5, 10)]          # reference bug 2: split over the subset itself


def test_partition_errors(b2):
    m = make_grm(b2, [])
    with pytest.raises(RuntimeError, match="No GPUs available"):
        m.distribute_workload(10)
    m = make_grm(b2, [0])
    with pytest.raises(ValueError):
        m.distribute_workload(0)
    with pytest.raises(ValueError, match="Unknown distribution strategy"):
        m.distribute_workload(10, "round_robin")


def test_validate_and_device_string(b2):
    m = make_grm(b2, [0, 1])
    assert not m.validate_gpu_index(-1) and not m.validate_gpu_index(2)
    with patch("torch.cuda.is_available", return_value=True), patch("torch.cuda.device_count", return_value=2):
        assert m.validate_gpu_index(1)
        assert m.get_safe_device_string(1) == "cuda:1"
        with pytest.raises(ValueError, match=r"Invalid GPU index: 5. Available GPUs: \[0, 1\]"):
            m.get_safe_device_string(5)
    assert "gpu_count=2" in str(m) and "gpu_configs=0" in repr(m)
    assert m.get_rank_info() == (0, 1)
    assert m.get_communicator() is None


def test_partition_even_helper(b2):
    assert b2.partition_even(301, 3) == [(0, 101), (101, 201), (201, 301)]
    assert b2.partition_even(10_000_000, 8)[7] == (8_750_000, 10_000_000)


# ---------------------------------------------------------------- EmbeddingDistributionManager
def mock_grm(b2, gpus=(0, 1)):
    g = Mock(spec=b2.GPUResourceManager)
    g.get_available_gpu_ids.return_value = list(gpus)
    g.validate_gpu_index.side_effect = lambda x: x in gpus
    g.get_safe_device_string.side_effect = lambda x: f"cuda:{x}"
    real = b2.GPUResourceManager()
    real.available_gpus = list(gpus)
    g.distribute_workload.side_effect = real.distribute_workload
    return g


def test_distribute_embeddings_host_logic(b2):
    g = mock_grm(b2)
    edm = b2.EmbeddingDistributionManager(g)
    emb = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4)
    with patch.object(torch.Tensor, "to", lambda self, *a, **k: self):
        dist = edm.distribute_embeddings(emb)
    assert [(p.gpu_id, p.start_index, p.end_index) for p in dist.parts] == [(0, 0, 4), (1, 4, 7)]
    assert torch.equal(dist.parts[1].tensor, emb[4:7])
    assert dist.total_size == 7 and dist.embedding_dim == 4
    summary = edm.get_distribution_summary(dist)
    assert summary["part_sizes"] == [4, 3] and summary["gpu_ids"] == [0, 1]
    assert edm.get_total_memory_usage(dist) == {0: 64, 1: 48}
    edm.cleanup_distribution(dist)
    assert edm.current_distribution is None


def test_distribute_embeddings_validation_messages(b2):
    edm = b2.EmbeddingDistributionManager(mock_grm(b2))
    with pytest.raises(TypeError, match="embeddings must be a torch.Tensor"):
        edm.distribute_embeddings([1, 2])
    with pytest.raises(ValueError, match="embeddings must be 2D tensor"):
        edm.distribute_embeddings(torch.zeros(3))
    with pytest.raises(ValueError, match="embeddings tensor cannot be empty"):
        edm.distribute_embeddings(torch.zeros(0, 4))
    with pytest.raises(ValueError, match="Target GPU 9 is not available"):
        edm.distribute_embeddings(torch.zeros(4, 4), target_gpus=[9])


def test_dataclass_validation(b2):
    t = torch.zeros(10, 4)
    with pytest.raises(ValueError, match="start_index must be non-negative"):
        b2.EmbeddingPart(0, t, -1, 9)
    with pytest.raises(ValueError, match="doesn't match index range"):
        b2.EmbeddingPart(0, t, 0, 20)
    a, b = b2.EmbeddingPart(0, t, 0, 10), b2.EmbeddingPart(1, t, 12, 22)
    with pytest.raises(ValueError, match="Gap or overlap detected"):
        b2.DistributedEmbeddings([a, b], 22, 4)
    with pytest.raises(ValueError, match="parts list cannot be empty"):
        b2.DistributedEmbeddings([], 1, 4)


# ---------------------------------------------------------------- IndexBuildingCoordinator
def sample_dist(b2):
    parts = [b2.EmbeddingPart(0, torch.randn(50, 8), 0, 50), b2.EmbeddingPart(1, torch.randn(50, 8), 50, 100)]
    return b2.DistributedEmbeddings(parts, 100, 8)


def ibc_grm(b2):
    g = Mock(spec=b2.GPUResourceManager)
    g.validate_gpu_index.return_value = True
    g.get_gpu_memory_info.return_value = {"allocated": 1, "reserved": 2, "total": 3, "free": 1}
    return g


def test_index_build_config_accepts_new_types(b2):
    for t in ("ivf_flat", "ivf_pq", "cagra", "brute_force", "flat"):
        b2.IndexBuildConfig(t, {})
    with pytest.raises(ValueError, match="index_type must be one of"):
        b2.IndexBuildConfig("hnsw", {})
    with pytest.raises(ValueError, match="timeout_seconds must be positive"):
        b2.IndexBuildConfig("flat", {}, timeout_seconds=0)


def test_simulated_build_without_cuda_and_bookkeeping(b2):
    ibc = b2.IndexBuildingCoordinator(ibc_grm(b2))
    cfg = b2.IndexBuildConfig("ivf_flat", {"n_lists": 4}, parallel_build=True, max_retries=0)
    res = ibc.build_indices_parallel(sample_dist(b2), cfg)
    assert res.success and res.successful_gpus == [0, 1]
    assert ibc.get_index_for_gpu(1) == {"type": "ivf_flat", "size": 50, "dim": 8}
    assert ibc.index_offsets == {0: 0, 1: 50}              # shard id offsets = start_index
    assert ibc.get_build_summary()["gpu_success_rates"] == {0: 1.0, 1: 1.0}
    assert not ibc.has_active_builds()
    ibc.cleanup_all_indices()
    assert ibc.built_indices == {}


def test_build_failure_and_retry(b2):
    g = ibc_grm(b2)
    calls = {"n": 0}

    def validate(gpu):
        if gpu == 0:
            return True
        calls["n"] += 1
        return calls["n"] > 1
    g.validate_gpu_index.side_effect = validate
    ibc = b2.IndexBuildingCoordinator(g)
    cfg = b2.IndexBuildConfig("flat", {}, parallel_build=False, max_retries=1)
    assert ibc.build_indices_parallel(sample_dist(b2), cfg).success
    g.validate_gpu_index.side_effect = lambda gpu: gpu == 0
    cfg0 = b2.IndexBuildConfig("flat", {}, parallel_build=False, max_retries=0)
    res = ibc.build_indices_parallel(sample_dist(b2), cfg0)
    assert not res.success and res.failed_gpus == [1] and list(ibc.built_indices) == [0]
    assert "Failed after 1 attempts" in res.build_results[1].error_message


# ---------------------------------------------------------------- SearchResultAggregator
def sra_fixture(b2):
    g = Mock(spec=b2.GPUResourceManager)
    g.validate_gpu_index.return_value = True
    g.get_safe_device_string.side_effect = lambda x: f"cuda:{x}"
    return b2.SearchResultAggregator(g), g


def sr(b2, d, i, gpu):
    d = np.asarray(d, np.float32)
    return b2.SearchResult(d, np.asarray(i, np.int64), gpu, 0.1, d.shape[1], d.shape[1])


def test_merge_known_answers_from_reference_tests(b2):
    agg, _ = sra_fixture(b2)
    fd, fi = agg.merge_search_results([sr(b2, [[1, 2, 3], [4, 5, 6]], [[10, 20, 30], [40, 50, 60]], 0)], 2)
    np.testing.assert_array_equal(fd, [[1, 2], [4, 5]])
    np.testing.assert_array_equal(fi, [[10, 20], [40, 50]])
    two = [sr(b2, [[2, 4], [6, 8]], [[20, 40], [60, 80]], 0), sr(b2, [[1, 3], [5, 7]], [[10, 30], [50, 70]], 1)]
    fd, fi = agg.merge_search_results(two, 3)
    np.testing.assert_array_equal(fd, [[1, 2, 3], [5, 6, 7]])
    np.testing.assert_array_equal(fi, [[10, 20, 30], [50, 60, 70]])
    fd, fi = b2.combine_search_results(two, 3, descending=True)
    np.testing.assert_array_equal(fi, [[40, 30, 20], [80, 70, 60]])
    with pytest.raises(ValueError, match="Cannot merge empty results list"):
        agg.merge_search_results([], 2)
    with pytest.raises(ValueError, match="has.*queries, expected"):
        agg.merge_search_results([two[0], sr(b2, [[1, 2]], [[1, 2]], 1)], 2)


def test_merge_ties_keep_lower_shard_first(b2):
    agg, _ = sra_fixture(b2)
    a, b = sr(b2, [[1, 1]], [[7, 8]], 0), sr(b2, [[1, 1]], [[3, 4]], 1)
    _, fi = agg.merge_search_results([a, b], 3)
    np.testing.assert_array_equal(fi, [[7, 8, 3]])


def test_search_result_validation_order(b2):
    with pytest.raises(ValueError, match="distances must be 2D array"):
        b2.SearchResult(np.zeros(3, np.float32), np.zeros(3, np.int64), 0, 0.0, 3, 3)
    with pytest.raises(ValueError, match="distances shape.*!= indices shape"):
        b2.SearchResult(np.zeros((2, 3), np.float32), np.zeros((2, 2), np.int64), 0, 0.0, 3, 3)
    with pytest.raises(ValueError, match="k_returned.*cannot exceed k_requested"):
        b2.SearchResult(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.int64), 0, 0.0, 2, 3)
    with pytest.raises(ValueError, match="k must be positive"):
        b2.SearchConfig(k=0)


def test_distributed_search_simulated_and_validation(b2):
    agg, g = sra_fixture(b2)
    cfg = b2.SearchConfig(k=3, parallel_search=False)
    with pytest.raises(ValueError, match="query must be a torch.Tensor"):
        agg.perform_distributed_search("q", {0: Mock()}, cfg)
    with pytest.raises(ValueError, match="query must be 2D tensor"):
        agg.perform_distributed_search(torch.zeros(4), {0: Mock()}, cfg)
    with pytest.raises(ValueError, match="query cannot be empty"):
        agg.perform_distributed_search(torch.zeros(0, 4), {0: Mock()}, cfg)
    with pytest.raises(ValueError, match="indices dictionary cannot be empty"):
        agg.perform_distributed_search(torch.zeros(2, 4), {}, cfg)
    with patch("search_result_aggregator.CUVS_AVAILABLE", False):
        for parallel in (False, True):
            res = agg.perform_distributed_search(torch.randn(2, 4), {0: Mock(), 1: Mock()},
                                                 b2.SearchConfig(k=3, parallel_search=parallel))
            assert res.final_distances.shape == (2, 3) and len(res.gpu_results) == 2
            assert (np.diff(res.final_distances, axis=1) >= 0).all()
    assert len(agg.get_search_history()) == 2
    g.validate_gpu_index.return_value = False
    with pytest.raises(ValueError, match="GPU 99 in indices is not available"):
        agg.perform_distributed_search(torch.zeros(2, 4), {99: Mock()}, cfg)


def test_filter_by_distance(b2):
    r = b2.filter_search_results_by_distance(sr(b2, [[1, 2, 3]], [[5, 6, 7]], 0), 2.0)
    np.testing.assert_array_equal(r.indices, [[5, 6, -1]])
    assert np.isinf(r.distances[0, 2])


def test_recall_formula(b2):
    assert b2.recall_at_k([1, 2, 3, 4], [2, 4, 9], 3) == pytest.approx(1 / 3)
    assert b2.RecallEvaluator.calculate_recall_at_k([1, 2], [], 2) == 0.0
    truth = np.array([[1, 2], [3, 4]])
    assert b2.RecallEvaluator.batch_recall(np.array([[2, 9], [4, 3]]), truth, 2) == 0.75


def test_load_embedding_parts_reference_disk_format(b2, tmp_path):
    """embeddings_{size}_part{i}.pt files (cuvs-2gpu-main.ipynb cells 10/12), uneven parts."""
    a, b_ = torch.randn(38, 6), torch.randn(37, 6)
    pa, pb = tmp_path / "embeddings_75_part0.pt", tmp_path / "embeddings_75_part1.pt"
    torch.save(a, pa); torch.save(b_, pb)
    edm = b2.EmbeddingDistributionManager(mock_grm(b2))
    with patch.object(torch.Tensor, "to", lambda self, *a, **k: self):
        dist = edm.load_embedding_parts([str(pa), str(pb)])
    assert [(p.gpu_id, p.start_index, p.end_index) for p in dist.parts] == [(0, 0, 38), (1, 38, 75)]
    assert torch.equal(dist.parts[1].tensor, b_) and dist.total_size == 75
    full = tmp_path / "embeddings_75.pt"
    torch.save(torch.cat([a, b_]), full)
    with patch.object(torch.Tensor, "to", lambda self, *a, **k: self):
        dist = edm.load_embedding_parts([str(full)])
    assert [(p.start_index, p.end_index) for p in dist.parts] == [(0, 38), (38, 75)]
    with pytest.raises(ValueError, match="paths cannot be empty"):
        edm.load_embedding_parts([])


def test_degraded_search_reports_the_missing_shard(b2):
    """A shard that fails: by default the search fails with the shard's own exception; with
    search_params['allow_partial'] the other shards answer and the result names the missing GPUs
    (the reference drops failed shards with a log line only, improved_multi_gpu_rag.py:261-263)."""
    agg, g = sra_fixture(b2)
    real = agg._search_single_gpu

    def flaky(gpu_id, index, query, k_local, params):
        if gpu_id == 1:
            raise RuntimeError("Xid 79: GPU has fallen off the bus")
        return real(gpu_id, index, query, k_local, params)

    with patch("search_result_aggregator.CUVS_AVAILABLE", False), patch.object(agg, "_search_single_gpu", flaky):
        for parallel in (False, True):
            with pytest.raises(RuntimeError, match="fallen off the bus"):
                agg.perform_distributed_search(torch.randn(2, 4), {0: Mock(), 1: Mock(), 2: Mock()},
                                               b2.SearchConfig(k=3, parallel_search=parallel))
            res = agg.perform_distributed_search(
                torch.randn(2, 4), {0: Mock(), 1: Mock(), 2: Mock()},
                b2.SearchConfig(k=3, parallel_search=parallel, search_params={"allow_partial": True}))
            assert res.missing_gpus == [1] and not res.complete
            assert "fallen off the bus" in res.shard_errors[1]
            assert [r.gpu_id for r in res.gpu_results] == [0, 2]
            assert res.final_distances.shape == (2, 3)
        with pytest.raises(RuntimeError, match="failed on every shard"):
            agg.perform_distributed_search(torch.randn(2, 4), {1: Mock()},
                                           b2.SearchConfig(k=3, search_params={"allow_partial": True}))
    # a healthy search is complete and lists nothing
    with patch("search_result_aggregator.CUVS_AVAILABLE", False):
        res = agg.perform_distributed_search(torch.randn(2, 4), {0: Mock()}, b2.SearchConfig(k=3))
    assert res.complete and res.missing_gpus == [] and res.shard_errors == {}
