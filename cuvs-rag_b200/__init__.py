"""b2vs — B200-native sharded vector search behind the cuVS-rag manager API.

The package directory is named ``cuvs-rag_b200`` (not importable by name), so it is loaded
either through the repo-root shim ``cuvs_rag_b200.py`` (``import cuvs_rag_b200``) or by putting
this directory on ``sys.path`` and importing the reference's flat module names
(``gpu_resource_manager``, ``embedding_distribution_manager``, ``index_building_coordinator``,
``search_result_aggregator``, ``improved_multi_gpu_rag``) exactly as the reference's scripts and tests do.  Both spellings
resolve to ONE module object each (registered in ``sys.modules`` under both names), so classes
compare equal whichever way they were imported.
"""
import importlib
import sys

__version__ = "0.1.0"


def _load(name):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = importlib.import_module("." + name, __name__)
        sys.modules[name] = mod
    sys.modules[__name__ + "." + name] = mod
    return mod


_native = _load("_native")
gpu_resource_manager = _load("gpu_resource_manager")
embedding_distribution_manager = _load("embedding_distribution_manager")
index_building_coordinator = _load("index_building_coordinator")
search_result_aggregator = _load("search_result_aggregator")
evaluation = _load("evaluation")
improved_multi_gpu_rag = _load("improved_multi_gpu_rag")
encoder_handoff = _load("encoder_handoff")

from _native import NativeIndex, merge_topk, kmeans_fit, pool_normalize, build as build_native  # noqa: E402
from encoder_handoff import QueryEncoderHandoff, embed_queries, last_token_pool  # noqa: E402
from gpu_resource_manager import GPUResourceManager, GPUConfig, MultiGPUConfig, partition_even  # noqa: E402
from embedding_distribution_manager import (  # noqa: E402
    EmbeddingDistributionManager, EmbeddingPart, DistributedEmbeddings)
from index_building_coordinator import (  # noqa: E402
    IndexBuildingCoordinator, IndexBuildConfig, IndexBuildResult, CoordinatedIndexBuild)
from search_result_aggregator import (  # noqa: E402
    SearchResultAggregator, SearchResult, AggregatedSearchResult, SearchConfig,
    combine_search_results, filter_search_results_by_distance)
from evaluation import RecallEvaluator, recall_at_k  # noqa: E402
from improved_multi_gpu_rag import (  # noqa: E402
    IndexType, ParallelIndexBuilder, ParallelSearchEngine, CUDAMemoryManager)

__all__ = [
    "NativeIndex", "merge_topk", "kmeans_fit", "build_native",
    "GPUResourceManager", "GPUConfig", "MultiGPUConfig", "partition_even",
    "EmbeddingDistributionManager", "EmbeddingPart", "DistributedEmbeddings",
    "IndexBuildingCoordinator", "IndexBuildConfig", "IndexBuildResult", "CoordinatedIndexBuild",
    "SearchResultAggregator", "SearchResult", "AggregatedSearchResult", "SearchConfig",
    "combine_search_results", "filter_search_results_by_distance",
    "RecallEvaluator", "recall_at_k",
    "IndexType", "ParallelIndexBuilder", "ParallelSearchEngine", "CUDAMemoryManager",
    "pool_normalize", "QueryEncoderHandoff", "embed_queries", "last_token_pool",
]
