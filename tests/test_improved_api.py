"""Second-generation API (improved_multi_gpu_rag.py drop-in): golden parity of the host-side pieces
on CPU, the index builder / search engine through the C ABI on a GPU."""
import json
import os

import numpy as np
import pytest
import torch

import cuvs_rag_b200 as b2
from cuvs_rag_b200 import improved_multi_gpu_rag as imp

GOLD = os.path.join(os.path.dirname(__file__), "golden", "recall.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_recall_evaluator_matches_reference_outputs(gold):
    for c in gold["recall_cases"]:
        ret, rel = np.array(c["retrieved"], dtype=np.int64), np.array(c["relevant"], dtype=np.int64)
        for k in c["k_values"]:
            assert imp.RecallEvaluator.calculate_recall_at_k(ret, rel, k) == pytest.approx(c["single"][str(k)])
        multi = imp.RecallEvaluator.evaluate_recall_multiple_k(ret, rel, c["k_values"])
        assert {str(k): v for k, v in multi.items()} == pytest.approx(c["multi"])


def test_synthetic_ground_truth_and_defaults_match_reference(gold):
    gt = imp.RecallEvaluator.generate_synthetic_ground_truth(5, 1000, 20)
    assert {str(k): v.tolist() for k, v in gt.items()} == gold["synthetic_ground_truth"]
    cfg = imp.SearchConfig()
    assert {"top_k": cfg.top_k, "search_batch_size": cfg.search_batch_size,
            "num_queries": cfg.num_queries, "enable_recall_eval": cfg.enable_recall_eval,
            "recall_k_values": cfg.recall_k_values} == gold["search_config"]
    assert {t.name: t.value for t in imp.IndexType} == gold["index_types"]
    assert imp.GPUConfig(3).device_str == "cuda:3"
    assert imp.SearchConfig(recall_k_values=[1, 2]).recall_k_values == [1, 2]


def test_module_surface_and_aliases():
    import sys
    for name in ("IndexType", "SearchConfig", "GPUConfig", "CUDAMemoryManager", "ParallelIndexBuilder",
                 "ParallelSearchEngine", "RecallEvaluator", "get_memory_stats", "print_memory_status"):
        assert hasattr(imp, name), name
    assert sys.modules["improved_multi_gpu_rag"] is imp          # flat import name, like the reference
    assert b2.ParallelSearchEngine is imp.ParallelSearchEngine
    stats = imp.get_memory_stats()
    assert "ram_gb" in stats and "cpu_percent" in stats


def test_builder_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    builder = imp.ParallelIndexBuilder(1)
    res = builder.build_indices_parallel([torch.randn(100, 16)], imp.IndexType.IVF_FLAT)
    assert res["success"] is False and res["failed_gpus"] == [0] and res["indexes"] == {}
    with pytest.raises(RuntimeError, match="CUDA is required"):
        builder.build_index_on_gpu(imp.GPUConfig(0), torch.randn(10, 8), imp.IndexType.FAISS_FLAT, {})


@pytest.mark.gpu
def test_golden_host_merge_through_the_gpu_merge_kernel(gold):
    """ParallelSearchEngine's host merge (concatenate, argsort, top_k) == b2vs_merge_topk."""
    for m in gold["merges"]:
        d = torch.tensor(m["d"], dtype=torch.float32, device="cuda")[:, None, :]
        i = torch.tensor(m["i"], dtype=torch.int64, device="cuda")[:, None, :]
        od, oi = b2.merge_topk(d, i, m["top_k"])
        assert od[0].cpu().tolist() == pytest.approx(m["out_d"])
        assert oi[0].cpu().tolist() == m["out_i"]


@pytest.mark.gpu
@pytest.mark.parametrize("itype,k", [("FAISS_FLAT", 2000), ("IVF_FLAT", 2000), ("IVF_FLAT", 10),
                                     ("IVF_PQ", 10), ("IVF_PQ", 2000)])
def test_builder_and_engine_end_to_end(itype, k):
    from oracle.exact import exact_knn
    g = torch.Generator().manual_seed(3)
    cent = torch.randn(30, 64, generator=g)
    x = cent[torch.randint(0, 30, (24000,), generator=g)] + 0.5 * torch.randn(24000, 64, generator=g)
    parts = [x]                                      # one GPU on the test box; ids stay global
    builder = imp.ParallelIndexBuilder(1)
    res = builder.build_indices_parallel(parts, imp.IndexType[itype],
                                         {"n_lists": 16, "pq_dim": 32, "kmeans_n_iters": 8})
    assert res["success"] and set(res["indexes"]) == {0} and res["avg_time"] > 0
    eng = imp.ParallelSearchEngine(res["indexes"], imp.IndexType[itype], imp.SearchConfig(top_k=k))
    queries = [x[j] + 0.05 * torch.randn(64, generator=g) for j in (5, 77, 1234, 20000)]
    out = eng.batch_search(queries)
    assert len(out) == 4 and all(d.shape == (k,) and i.shape == (k,) for d, i in out)
    d1, i1 = eng.parallel_search(queries[1])
    assert d1.shape == (k,) and np.array_equal(i1, out[1][1])
    sd, si = eng.search_on_gpu(0, res["indexes"][0], queries[0], 5)
    assert isinstance(sd, np.ndarray) and sd.shape == (1, 5) and si.dtype == np.int64
    for (d, i), j in zip(out, (5, 77, 1234, 20000)):
        assert i[0] == j and bool(np.all(np.diff(d) >= -1e-4))
    if itype == "FAISS_FLAT":
        _, ti = exact_knn(x, torch.stack(queries), 100, "sqeuclidean")
        rec = imp.RecallEvaluator.evaluate_recall_multiple_k(out[2][1], ti[2].numpy(), [10, 100, 2000])
        assert rec[100] > 0.99 and rec[2000] > 0.99
