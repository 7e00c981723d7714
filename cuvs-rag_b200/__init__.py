"""b2vs — B200-native sharded vector search behind the cuVS-rag manager API.

The package directory is named ``cuvs-rag_b200`` (not importable by name), so it is loaded
either through the repo-root shim ``cuvs_rag_b200.py`` (``import cuvs_rag_b200``) or by putting
this directory on ``sys.path`` and importing the reference's flat module names
(``gpu_resource_manager``, ``embedding_distribution_manager``, ``index_building_coordinator``,
``search_result_aggregator``) exactly as the reference's scripts and tests do.
"""
from . import _native  # noqa: F401
from ._native import NativeIndex, merge_topk, kmeans_fit, build as build_native  # noqa: F401

__all__ = ["NativeIndex", "merge_topk", "kmeans_fit", "build_native"]
__version__ = "0.1.0"
